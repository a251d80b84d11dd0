// Mixed-radix frame lengths of the small-frame encoder (enc2048.cu: encs_kernel), 64-bit arithmetic.
#define FRAD_TU_MIXED 64
#include "enc2048.cu"
