// engine_launch.h - host-side launchers of the six fused frame-step kernels.  Each family is instantiated in its own
// translation unit (engine_<name>.cu) so that the library builds in parallel; cabi.cu only sees these declarations.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace mot {
struct BtArgs; struct SortArgs; struct OcArgs; struct BotArgs; struct SsArgs; struct BoostArgs;
// *_prepare: opt the kernel of `shape` into `smem` bytes of dynamic shared memory; *_launch: one CTA per stream
// ByteTrack picks its CTA width from the stream count: 512 threads x 2 CTAs per SM, or one 1024-thread CTA per SM
int bt_threads(int shape, int n_streams, int n_sms);
cudaError_t bt_prepare(int shape, size_t smem, int threads);
void bt_launch(int shape, int grid, size_t smem, cudaStream_t st, const BtArgs& a, int threads);
cudaError_t sort_prepare(int shape, size_t smem);
void sort_launch(int shape, int grid, size_t smem, cudaStream_t st, const SortArgs& a);
cudaError_t oc_prepare(int shape, size_t smem);
void oc_launch(int shape, int grid, size_t smem, cudaStream_t st, const OcArgs& a);
cudaError_t oc_centroid_prepare(int shape, size_t smem);
void oc_centroid_launch(int shape, int grid, size_t smem, cudaStream_t st, const OcArgs& a);
cudaError_t deepoc_prepare(int shape, size_t smem);
void deepoc_launch(int shape, int grid, size_t smem, cudaStream_t st, const OcArgs& a);
cudaError_t boost_prepare(int shape, size_t smem);
void boost_launch(int shape, int grid, size_t smem, cudaStream_t st, const BoostArgs& a);
cudaError_t bot_prepare(int shape, size_t smem);
void bot_launch(int shape, int grid, size_t smem, cudaStream_t st, const BotArgs& a);
cudaError_t ss_prepare(int shape, size_t smem);
void ss_launch(int shape, int grid, size_t smem, cudaStream_t st, const SsArgs& a);
}  // namespace mot
