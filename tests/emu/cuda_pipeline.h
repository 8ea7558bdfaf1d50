// cuda_pipeline.h of the CPU kernel emulator -- TEST INFRASTRUCTURE ONLY (tests/emu/README.md).
// cp.async (LDGSTS) primitives as immediate copies: a fiber sees its own copies at once, other fibers after the
// next barrier, which is all the kernels rely on.
#pragma once
#include <string.h>
static inline void __pipeline_memcpy_async(void *dst_shared, const void *src_global, size_t size) { memcpy(dst_shared, src_global, size); }
static inline void __pipeline_commit() {}
static inline void __pipeline_wait_prior(int) {}
