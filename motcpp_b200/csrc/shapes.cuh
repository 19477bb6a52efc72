// shapes.cuh - the (track capacity, detections per frame, candidate-edge buffer) shapes the fused ByteTrack / SORT kernels
// are compiled for; BoT-SORT / StrongSORT / OC-SORT keep their own tables next to their kernels.
#pragma once

namespace mot {

struct BtShape { int cap, d_max, e_cap; };
constexpr BtShape kBtShapes[] = {{256, 64, 1024}, {1536, 512, 4096}, {2048, 512, 4096}, {3072, 1024, 4096}};
constexpr int kNumBtShapes = sizeof(kBtShapes) / sizeof(kBtShapes[0]);

}  // namespace mot
