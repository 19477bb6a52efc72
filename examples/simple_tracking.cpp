// The reference's examples/simple_tracking.cpp loop, on the B200 engine: construct a tracker with
// the reference's positional arguments, call update(dets, img) per frame.
//   g++ -std=c++17 -Iinclude examples/simple_tracking.cpp -Lmotcpp_b200 -lmotb200 -Wl,-rpath,$PWD/motcpp_b200
#include <cstdio>
#include <motcpp_b200/trackers.hpp>

int main() {
    try {
        motcpp_b200::ByteTrack tracker(0.3f, 30, 50, 3, 0.3f, false, 80, "iou", false, 0.1f, 0.45f, 0.8f, 30, 30);
        // the other front-ends, with the argument lists tools/motcpp_eval.cpp passes them
        motcpp_b200::Sort sort(0.3f, 1, 50, 3, 0.3f);
        motcpp_b200::OCSort ocsort(0.2f, 30, 50, 3, 0.3f, false, 80, "iou", false, 0.1f, 3, 0.2f, false, 0.01f, 0.0001f);
        motcpp_b200::BotSort botsort("", false, false, 0.3f, 30, 50, 3, 0.3f, false, 80, "iou", false, 0.6f, 0.1f, 0.7f, 30,
                                     0.8f, 0.5f, 0.25f, "none", 30, false, true, /*emb_dim=*/4);
        motcpp_b200::StrongSORT strongsort("", false, false, 0.3f, 30, 50, 3, 0.3f, false, 80, "iou", false, 0.1f, 0.2f, 0.7f, 3,
                                           100, 0.98f, 0.9f, /*emb_dim=*/4, /*track_capacity=*/256, /*max_dets=*/64);
        motcpp_b200::DeepOCSort deepocsort("", false, false, 0.3f, 30, 50, 3, 0.3f, false, 80, "iou", false, 3, 0.2f, 0.5f, 0.95f,
                                           0.5f, false, /*cmc_off=*/true, false, 0.01f, 0.0001f, /*emb_dim=*/4,
                                           /*track_capacity=*/256, /*max_dets=*/64);
        motcpp_b200::BoostTrackTracker boosttrack("", false, false, 0.6f, 60, 50, 3, 0.3f, false, 80, "iou", false, /*use_ecc=*/false,
                                                  10, 1.6f, "ecc", 0.5f, 0.25f, 0.25f, true, true, 0.65f, false, false, false, false,
                                                  /*with_reid=*/false, /*track_capacity=*/256, /*max_dets=*/64);
        cv::Mat img(480, 640);
        Eigen::MatrixXf dets(2, 6);
        const float rows[2][6] = {{100, 100, 200, 200, 0.9f, 0}, {300, 300, 400, 420, 0.8f, 0}};
        for (int frame = 0; frame < 3; ++frame) {
            for (int i = 0; i < 2; ++i)
                for (int c = 0; c < 6; ++c) dets(i, c) = rows[i][c] + (c < 4 ? 2.0f * frame : 0.0f);
            Eigen::MatrixXf embs(2, 4);
            for (int i = 0; i < 2; ++i)
                for (int c = 0; c < 4; ++c) embs(i, c) = (c == i) ? 1.0f : 0.1f;
            std::printf("frame %d rows: sort %ld ocsort %ld botsort %ld\n", frame, (long)sort.update(dets, img).rows(),
                        (long)ocsort.update(dets, img).rows(), (long)botsort.update(dets, img, embs).rows());
            std::printf("frame %d strongsort rows %ld\n", frame, (long)strongsort.update(dets, img, embs).rows());
            std::printf("frame %d deepocsort rows %ld\n", frame, (long)deepocsort.update(dets, img, embs).rows());
            std::printf("frame %d boosttrack rows %ld\n", frame, (long)boosttrack.update(dets, img).rows());
            const Eigen::MatrixXf tracks = tracker.update(dets, img);
            for (long i = 0; i < tracks.rows(); ++i)
                std::printf("frame %d id %d box %.1f %.1f %.1f %.1f conf %.2f\n", frame, (int)tracks(i, 4), tracks(i, 0),
                            tracks(i, 1), tracks(i, 2), tracks(i, 3), tracks(i, 5));
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
