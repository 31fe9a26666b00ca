// placeholder (replaced by the full standard model)
#pragma once
#include "glc_model_box.cuh"
namespace glc {
struct ModelStandard {
    static __device__ __forceinline__ uint32_t active_mask(int) { return 0; }
    static __device__ __forceinline__ void solve_analytics(NodeCtx &, double) {}
    static __device__ __forceinline__ void scales(const NodeCtx &, const double (&)[NY], double (&)[NY]) {}
    static __device__ __forceinline__ int rates(NodeCtx &, double, const double (&)[NY], double (&)[NY]) { return 0; }
    static __device__ __forceinline__ int post_step(NodeCtx &, double (&)[NY]) { return 0; }
    static __device__ __forceinline__ void pre_evolve(NodeCtx &, double (&)[NY]) {}
    static __device__ __forceinline__ void post_evolve(NodeCtx &, double (&)[NY]) {}
};
}
