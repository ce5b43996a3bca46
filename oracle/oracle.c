/*
 * oracle.c — TEST INFRASTRUCTURE ONLY.  See oracle_impl.h for scope, provenance and the
 * "parity unpinned" statement.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the library built from this file.
 *
 * Build: make -C oracle   ->  oracle/_build/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define SUF(name) CAT(name, _f32)
#include "oracle_impl.h"
#undef REAL
#undef SUF
#undef NEG_INF

#define oracle_hash32 oracle_hash32_d
#define oracle_drop_keep oracle_drop_keep_d
#define REAL double
#define SUF(name) CAT(name, _f64)
#include "oracle_impl.h"
#undef REAL
#undef SUF

int oracle_abi_version(void) { return 1; }
