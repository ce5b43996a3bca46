"""ctc_asr_b200 — B200-native (sm_100a) drop-in for the hot path of mdangschat/ctc-asr:
`CTCModel.inference_fn` / `loss_fn` / `decode_fn` (asr/model.py:123-309) over libctcasr.so."""
from .params import FLAGS, ModelConfig  # noqa: F401


def __getattr__(name):          # keep `import ctc_asr_b200` torch-free for config-only users
    if name == "CTCModel":
        from .model import CTCModel
        return CTCModel
    raise AttributeError(name)
