// Force-included (-include) ahead of the reference's project_cloud.cu when building oracle/_ref.
// The reference cudaMalloc's its per-resolution buffers and never clears them; at resolutions
// that are not multiples of 16 (1920x1080) the tail of the z-buffer / image / tensor is never
// written by any kernel, so its content is whatever cudaMalloc returned.
// BASELINE.md defines parity "against the oracle run with zero-initialised buffers": this prelude
// makes that deterministic by routing the reference's cudaMalloc CALL SITES (not the runtime's
// declarations, which are already parsed by the time the macro exists) through a zeroing wrapper.
#pragma once
#include <cuda_runtime.h>
#include <torch/script.h>
#include <torch/cuda.h>
extern "C" cudaError_t rtr_ref_zero_malloc(void** p, size_t bytes);
#define cudaMalloc(p, s) rtr_ref_zero_malloc((void**)(p), (s))
