// kernels_cosine.cuh - embedding cosine-distance cost matrix (tcgen05 tensor-core contraction).
// Placeholder until the tcgen05 kernel lands: fails loudly instead of computing anything.
#pragma once
#include <string>
#include "simt.cuh"
namespace mot {
inline int launch_cosine(const float*, int, const float*, int, int, float*, int, cudaStream_t, std::string& err) {
    err = "mot_cost_cosine: tcgen05 kernel not built into this library yet";
    return 6;   // MOT_ERR_UNSUPPORTED
}
}  // namespace mot
