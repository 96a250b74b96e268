"""Node kinds and renderer flags of include/friendship_b200.h, as plain numbers: the workload builders must not import
the product package (bench.py's reference arm runs them against the CPU oracle alone).  tests/test_cabi_symbols.py
checks them against libfriendship_b200._cabi."""
# the seven primitives of reference src/routing/effect.rs:86-112
KIND_DELAY = 0
KIND_F32CONSTANT = 1
KIND_SUM2 = 2
KIND_MULTIPLY = 3
KIND_DIVIDE = 4
KIND_MODULO = 5
KIND_MINIMUM = 6
KIND_EFFECT = 16          # nested effect (reference effect.rs:79-83)
# extension nodes (not in the reference)
KIND_OSCBANK = 32
KIND_DIRECTFORM = 33
KIND_FBDELAY = 34

FLAG_SPARKLE_DELAY = 1
FLAG_NO_JIT = 2
FLAG_JIT_EAGER = 4
FLAG_NO_CHAIN_FUSION = 8
FLAG_NO_EXCITER_FUSION = 16
FLAG_SPARKLE_MIN = 32
FLAG_NO_TENSOR_OSC = 64
