"""Alphabet <-> integer labels, same API and numbering as the reference (asr/labels.py:11-59).

ids: 0 = unused / padding (decodes to ''), 1 = ' ', 2..27 = 'a'..'z', 28 = CTC blank
(tf.nn.ctc_loss puts the blank at num_classes - 1; asr/labels.py:6).
"""

ALPHABET = " abcdefghijklmnopqrstuvwxyz"
PAD_ID = 0


def num_classes():
    """27 characters + the unused id 0 + the CTC blank = 29 (asr/labels.py:53-59)."""
    return len(ALPHABET) + 2


def blank_id():
    return num_classes() - 1


def ctoi(char):
    """Character -> integer label; raises ValueError like asr/labels.py:29-35."""
    if len(char) != 1:
        raise ValueError('"{}" is not a valid character.'.format(char))
    pos = ALPHABET.find(char)
    if pos < 0:
        raise ValueError("Invalid input character '{}'.".format(char))
    return pos + 1


def itoc(integer):
    """Integer label -> character; 0 maps to '' (asr/labels.py:13), out of range raises."""
    if not 0 <= integer < num_classes():
        raise ValueError("Integer label ({}) out of range.".format(integer))
    if integer == PAD_ID or integer == blank_id():
        if integer == PAD_ID:
            return ""
        raise KeyError(integer)     # the reference's dict has no entry for the blank either
    return ALPHABET[integer - 1]


def text_to_ids(text):
    return [ctoi(c) for c in text]


def ids_to_text(ids):
    return "".join(itoc(int(i)) for i in ids if 0 <= int(i) < blank_id())
