/* PRIVATE profiling hooks of libcwfa_b200.so -- not part of the drop-in boundary (include/cwfa_b200.h).
 * Used only by scripts/trace_*.py to record per-CTA globaltimer stamps inside the tcgen05 kernels. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
/* device buffer of 8 uint64 per CTA (NULL = off): conv_tc_kernel stamps its pipeline phases */
int cwfa_tc_set_debug_buffer(void* buf);
/* [cta][8 tiles][8 stamps] uint64 (NULL = off): resblock_tc_kernel per-tile timeline */
int cwfa_resblock_set_debug_buffer(void* buf);
/* 8 uint64 (NULL = off): stencil3d_tc_kernel CTA (0,0): cycles in [x load, im2col, GEMM1, epilogue 1, GEMM2, epilogue 2, gather], pixel-rows */
int cwfa_stencil_set_debug_buffer(void* buf);
#ifdef __cplusplus
}
#endif
