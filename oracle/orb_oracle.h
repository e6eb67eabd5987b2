// oracle/orb_oracle.h -- TEST INFRASTRUCTURE ONLY (the checker, never the product path).
//
// CPU restatement of the reference's ORB extraction path, function by function:
//   ORBextractor::ORBextractor        /root/reference/src/ORBextractor.cc:468-571
//   IC_Angle / computeOrientation     :91-138, :580-591
//   computeOrbDescriptor              :150-203 (pattern :206-464)
//   ExtractorNode::DivideNode         :602-674
//   compareNodes / DistributeOctTree  :676-697, :711-1057
//   ComputeKeyPointsOctTree           :1061-1208
//   computeDescriptors / operator()   :1534-1547, :1557-1682
//   ComputePyramid                    :1687-1740
// OpenCV primitives come from cvprims.* (restated, pinned to cv2 4.13.0).  Parity status:
// the reference holds no tests or golden vectors for this path, so
// the restatement is pinned against (i) cv2 4.13.0 for every primitive, (ii) the
// reference's own ORBextractor.cc compiled verbatim against oracle/cvshim (oracle/_ref),
// see tests/test_oracle_vs_ref.py.
#pragma once
#include <cstdint>
#include <vector>

namespace orb_oracle {

struct OrbKp {  // byte-compatible with cv::KeyPoint (28 B)
    float x, y, size, angle, response;
    int octave, class_id;
};

struct Cand {  // one FAST candidate in window coordinates (origin = minBorder)
    int x, y, score;
};

struct Level {
    int w = 0, h = 0;            // ROI size
    int step = 0;                // padded row pitch = w + 38
    std::vector<uint8_t> padded; // (h+38) x (w+38), ROI at (19,19)
    std::vector<uint8_t> blurred;  // h x w
    std::vector<Cand> cands;       // emission order (cells row-major, FAST row-major)
    std::vector<uint8_t> cell_retry;  // per cell: 0 = iniTh hit, 1 = minTh used, 2 = skipped
    std::vector<OrbKp> kps;        // retained, list order, level coordinates, with angle
    const uint8_t* roi() const { return padded.data() + 19 * step + 19; }
};

class Extractor {
   public:
    Extractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    // Returns monoIndex, or -1 for an empty image (reference :1561-1562).
    int extract(const uint8_t* img, int rows, int cols, size_t step, int lap0, int lap1,
                std::vector<OrbKp>& kps, std::vector<uint8_t>& desc);
    void compute_pyramid(const uint8_t* img, int rows, int cols, size_t step);
    void compute_keypoints_octtree();

    int nfeatures, nlevels, iniThFAST, minThFAST;
    double scaleFactor;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel, umax;
    std::vector<Level> levels;
};

// DistributeOctTree on candidates given in window coordinates; returns indices (into
// `cands`) of the retained keypoints in list order.
std::vector<int> distribute_octtree(const std::vector<Cand>& cands, int minX, int maxX, int minY,
                                    int maxY, int N);

float ic_angle(const uint8_t* center, int step, const std::vector<int>& umax);
void orb_descriptor(const uint8_t* center, int step, float angle_deg, uint8_t* desc);
extern const int8_t kPattern[1024];

}  // namespace orb_oracle
