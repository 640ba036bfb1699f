// Fast mode: float32, wide-node traversal, FMA contraction allowed.
#define MCRT_REAL float
#include "kernels_impl.cuh"
