// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY: flat C entry points so that tests/
// and bench.py's cpu_baseline leg can drive the CPU oracle through ctypes.
#include <cstring>
#include <vector>

#include "../include/orbfe.h"
#include "cvprims.h"
#include "match_oracle.h"
#include "bow_oracle.h"
#include "orb_oracle.h"

using orb_oracle::Extractor;
using orb_oracle::OrbKp;

extern "C" {

// ---- OpenCV primitives ------------------------------------------------------------------
void oracle_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstep, uint8_t* dst, int dw,
                             int dh, int dstep) {
    cvp::resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void oracle_border101(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep,
                      int pad) {
    cvp::copy_make_border_reflect101(src, w, h, sstep, dst, dstep, pad, pad, pad, pad);
}
int oracle_fast(const uint8_t* img, int w, int h, int step, int th, int nms, int* xys, int cap) {
    std::vector<cvp::FastKP> out;
    cvp::fast9_16(img, w, h, step, th, nms != 0, out);
    int n = (int)out.size();
    for (int i = 0; i < n && i < cap; i++) {
        xys[3 * i] = out[i].x; xys[3 * i + 1] = out[i].y; xys[3 * i + 2] = out[i].score;
    }
    return n;
}
void oracle_blur7(const uint8_t* src, int w, int h, int sstep, uint8_t* dst, int dstep) {
    cvp::gaussian_blur_7x7_s2(src, w, h, sstep, dst, dstep);
}
void oracle_fast_atan2(const float* y, const float* x, float* out, int n) {
    for (int i = 0; i < n; i++) out[i] = cvp::fast_atan2(y[i], x[i]);
}
int oracle_hamming(const uint8_t* a, const uint8_t* b) { return cvp::hamming256(a, b); }
void oracle_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx, int* dist) {
    cvp::bf_knn2(q, nq, t, nt, idx, dist);
}

// ---- extractor --------------------------------------------------------------------------
void* oracle_extractor_create(int nf, float sf, int nl, int ini, int mn) {
    return new Extractor(nf, sf, nl, ini, mn);
}
void oracle_extractor_destroy(void* h) { delete (Extractor*)h; }

int oracle_extract(void* h, const uint8_t* img, int rows, int cols, int step, int lap0, int lap1,
                   OrbKp* kps, uint8_t* desc, int cap, int* n_out) {
    Extractor* e = (Extractor*)h;
    std::vector<OrbKp> k;
    std::vector<uint8_t> d;
    int mono = e->extract(img, rows, cols, step, lap0, lap1, k, d);
    if (mono < 0) { *n_out = 0; return -1; }
    *n_out = (int)k.size();
    int n = (int)k.size() < cap ? (int)k.size() : cap;
    if (n) {
        memcpy(kps, k.data(), sizeof(OrbKp) * n);
        memcpy(desc, d.data(), 32 * (size_t)n);
    }
    return mono;
}
void oracle_tables(void* h, float* scale, float* inv, float* sig2, float* invsig2, int* nfeat,
                   int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; i++) {
        scale[i] = e->mvScaleFactor[i]; inv[i] = e->mvInvScaleFactor[i];
        sig2[i] = e->mvLevelSigma2[i]; invsig2[i] = e->mvInvLevelSigma2[i];
        nfeat[i] = e->mnFeaturesPerLevel[i];
    }
    for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}
void oracle_level_dims(void* h, int lvl, int* w, int* hh, int* ncand, int* nkp, int* ncells) {
    const orb_oracle::Level& L = ((Extractor*)h)->levels[lvl];
    *w = L.w; *hh = L.h; *ncand = (int)L.cands.size(); *nkp = (int)L.kps.size();
    *ncells = (int)L.cell_retry.size();
}
// padded: (h+38)*(w+38) bytes; blurred: h*w bytes (may be null)
void oracle_level_images(void* h, int lvl, uint8_t* padded, uint8_t* blurred) {
    const orb_oracle::Level& L = ((Extractor*)h)->levels[lvl];
    if (padded) memcpy(padded, L.padded.data(), L.padded.size());
    if (blurred && !L.blurred.empty()) memcpy(blurred, L.blurred.data(), L.blurred.size());
}
void oracle_level_lists(void* h, int lvl, int* cands_xys, uint8_t* cell_retry, OrbKp* kps) {
    const orb_oracle::Level& L = ((Extractor*)h)->levels[lvl];
    if (cands_xys)
        for (size_t i = 0; i < L.cands.size(); i++) {
            cands_xys[3 * i] = L.cands[i].x; cands_xys[3 * i + 1] = L.cands[i].y;
            cands_xys[3 * i + 2] = L.cands[i].score;
        }
    if (cell_retry && !L.cell_retry.empty()) memcpy(cell_retry, L.cell_retry.data(), L.cell_retry.size());
    if (kps && !L.kps.empty()) memcpy(kps, L.kps.data(), sizeof(OrbKp) * L.kps.size());
}
// Stand-alone octree: cands (x,y,score) in window coordinates -> retained indices.
int oracle_octree(const int* xys, int n, int minX, int maxX, int minY, int maxY, int N, int* keep,
                  int cap) {
    std::vector<orb_oracle::Cand> c(n);
    for (int i = 0; i < n; i++) c[i] = {xys[3 * i], xys[3 * i + 1], xys[3 * i + 2]};
    std::vector<int> r = orb_oracle::distribute_octtree(c, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) keep[i] = r[i];
    return (int)r.size();
}
float oracle_ic_angle(const uint8_t* img, int step, int x, int y) {
    static Extractor e(1000, 1.2f, 8, 20, 7);
    return orb_oracle::ic_angle(img + (size_t)y * step + x, step, e.umax);
}
void oracle_descriptor(const uint8_t* img, int step, int x, int y, float angle, uint8_t* desc) {
    orb_oracle::orb_descriptor(img + (size_t)y * step + x, step, angle, desc);
}


// ---- matchers (struct layouts shared with include/orbfe.h) -----------------------------
int oracle_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return match_oracle::descriptor_distance(a, b);
}

int oracle_search_by_projection(const OrbfeFrameView* fv, const OrbfeProjPoints* pp,
                                const OrbfeSearchParams* prm, const float* scale_factors,
                                int nlevels, const uint8_t* claimed, int32_t* assigned,
                                int32_t* best_idx, int32_t* best_dist) {
    match_oracle::FrameView F;
    F.N = fv->n; F.keys = (const OrbKp*)fv->keys; F.uright = fv->uright; F.desc = fv->desc;
    F.minX = fv->min_x; F.minY = fv->min_y; F.maxX = fv->max_x; F.maxY = fv->max_y;
    F.gridWInv = fv->grid_w_inv; F.gridHInv = fv->grid_h_inv;
    F.scaleFactors = scale_factors; F.nlevels = nlevels;
    F.assign_features_to_grid();
    std::vector<match_oracle::ProjPoint> pts(pp->m);
    for (int j = 0; j < pp->m; j++) {
        match_oracle::ProjPoint& p = pts[j];
        p.u = pp->u[j]; p.v = pp->v[j]; p.ur = pp->ur ? pp->ur[j] : 0.f;
        p.radius = pp->radius[j]; p.minLevel = pp->min_level[j]; p.maxLevel = pp->max_level[j];
        p.angle = pp->angle ? pp->angle[j] : 0.f;
        p.valid = pp->valid ? pp->valid[j] : 1; p.blocks = pp->blocks ? pp->blocks[j] : 1;
    }
    match_oracle::SearchParams sp{prm->mode, prm->th_accept, prm->nnratio, prm->check_orientation};
    std::vector<int> bi(pp->m), bd(pp->m);
    int n = match_oracle::search_by_projection(F, pts, pp->desc, sp, claimed, assigned, bi.data(), bd.data());
    if (best_idx) memcpy(best_idx, bi.data(), sizeof(int) * pp->m);
    if (best_dist) memcpy(best_dist, bd.data(), sizeof(int) * pp->m);
    return n;
}

static void fill_view(match_oracle::FrameView& F, const OrbfeFrameView* fv) {
    F.N = fv->n; F.keys = (const OrbKp*)fv->keys; F.uright = fv->uright; F.desc = fv->desc;
    F.minX = fv->min_x; F.minY = fv->min_y; F.maxX = fv->max_x; F.maxY = fv->max_y;
    F.gridWInv = fv->grid_w_inv; F.gridHInv = fv->grid_h_inv;
    F.assign_features_to_grid();
}
static std::vector<match_oracle::ProjPoint> fill_pts(const OrbfeProjPoints* pp, const OrbfeProjPoints* shared) {
    std::vector<match_oracle::ProjPoint> pts(pp->m);
    for (int j = 0; j < pp->m; j++) {
        match_oracle::ProjPoint& p = pts[j];
        p.u = pp->u[j]; p.v = pp->v[j]; p.ur = 0.f;
        p.radius = pp->radius[j]; p.minLevel = pp->min_level[j]; p.maxLevel = pp->max_level[j];
        p.angle = shared->angle ? shared->angle[j] : 0.f;
        p.valid = pp->valid ? pp->valid[j] : 1; p.blocks = shared->blocks ? shared->blocks[j] : 1;
    }
    return pts;
}
int oracle_search_by_projection_fisheye(const OrbfeFrameView* fl, const OrbfeFrameView* fr, const int32_t* l2r,
                                        const int32_t* r2l, const OrbfeProjPoints* pl, const OrbfeProjPoints* pr,
                                        const OrbfeSearchParams* prm, const uint8_t* claimed, int32_t* assigned,
                                        int32_t* best_l, int32_t* best_r) {
    match_oracle::FrameView FL, FR;
    fill_view(FL, fl);
    fill_view(FR, fr);
    std::vector<match_oracle::ProjPoint> ptsL = fill_pts(pl, pl), ptsR = fill_pts(pr, pl);
    match_oracle::SearchParams sp{prm->mode, prm->th_accept, prm->nnratio, prm->check_orientation};
    return match_oracle::search_by_projection_fisheye(FL, FR, l2r, r2l, ptsL, ptsR, pl->desc, sp, claimed, assigned,
                                                      best_l, best_r);
}

int oracle_search_for_initialization(const OrbfeFrameView* f1, const OrbfeFrameView* f2, float* prev, int windowSize,
                                     float nnratio, int checkOri, int32_t* matches12) {
    match_oracle::FrameView F1, F2;
    F1.N = f1->n; F1.keys = (const OrbKp*)f1->keys; F1.desc = f1->desc;
    fill_view(F2, f2);
    return match_oracle::search_for_initialization(F1, F2, prev, windowSize, nnratio, checkOri != 0, matches12);
}

void oracle_search_window(const OrbfeFrameView* kf, const OrbfeProjPoints* pp, int thAccept, int fuseGate,
                          const float* invLevelSigma2, int nlevels, int32_t* best_idx, int32_t* best_dist) {
    match_oracle::FrameView F;
    fill_view(F, kf);
    std::vector<match_oracle::ProjPoint> pts = fill_pts(pp, pp);
    for (int j = 0; j < pp->m; j++) pts[j].ur = pp->ur ? pp->ur[j] : 0.f;
    const match_oracle::WindowParams wp{thAccept, fuseGate, invLevelSigma2, nlevels};
    match_oracle::search_window(F, pts, pp->desc, wp, best_idx, best_dist);
}

int oracle_search_by_sim3(const OrbfeFrameView* kf1, const OrbfeFrameView* kf2, const OrbfeProjPoints* p12,
                          const OrbfeProjPoints* p21, int thAccept, int32_t* match12) {
    match_oracle::FrameView F1, F2;
    fill_view(F1, kf1);
    fill_view(F2, kf2);
    std::vector<match_oracle::ProjPoint> a = fill_pts(p12, p12), b = fill_pts(p21, p21);
    return match_oracle::search_by_sim3(F1, F2, a, p12->desc, b, p21->desc, thAccept, match12);
}

// Grid query tap: indices returned by GetFeaturesInArea, in the reference's order.
int oracle_features_in_area(const OrbfeFrameView* fv, float x, float y, float r, int minLevel,
                            int maxLevel, int32_t* out, int cap) {
    match_oracle::FrameView F;
    F.N = fv->n; F.keys = (const OrbKp*)fv->keys;
    F.minX = fv->min_x; F.minY = fv->min_y; F.maxX = fv->max_x; F.maxY = fv->max_y;
    F.gridWInv = fv->grid_w_inv; F.gridHInv = fv->grid_h_inv;
    F.assign_features_to_grid();
    std::vector<int> v = F.features_in_area(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

// Stereo: both extractors must hold the pyramids of the pair (oracle_extract called on each).
void oracle_stereo_match(void* hl, void* hr, const OrbKp* kl, const uint8_t* dl, int nl,
                         const OrbKp* kr, const uint8_t* dr, int nr, float mbf, float mb,
                         float* uright, float* depth) {
    Extractor* L = (Extractor*)hl; Extractor* R = (Extractor*)hr;
    std::vector<match_oracle::PyrLevelView> pl(L->nlevels), pr(R->nlevels);
    for (int i = 0; i < L->nlevels; i++) {
        pl[i] = {L->levels[i].roi(), L->levels[i].w, L->levels[i].h, L->levels[i].step};
        pr[i] = {R->levels[i].roi(), R->levels[i].w, R->levels[i].h, R->levels[i].step};
    }
    match_oracle::compute_stereo_matches(kl, dl, nl, kr, dr, nr, pl.data(), pr.data(),
                                         L->mvScaleFactor.data(), L->mvInvScaleFactor.data(), mbf,
                                         mb, uright, depth);
}

void oracle_fisheye_matches(const uint8_t* q, int nq, const uint8_t* t, int nt, int* match,
                            int* idx2, int* dist2) {
    match_oracle::fisheye_ratio_matches(q, nq, t, nt, match, idx2, dist2);
}


void oracle_cvt_gray(const uint8_t* src, int w, int h, int sstep, int channels, int rgb, uint8_t* dst, int dstep) {
    cvp::cvt_gray_u8(src, w, h, sstep, channels, rgb != 0, dst, dstep);
}
void oracle_remap_linear(const uint8_t* src, int sw, int sh, int sstep, const float* mapx, const float* mapy, int dw, int dh,
                         uint8_t* dst, int dstep) {
    cvp::remap_linear_u8(src, sw, sh, sstep, mapx, mapy, dw, dh, dst, dstep);
}
void oracle_undistort_points(const float* xy, int n, double fx, double fy, double cx, double cy, const double* dist, int ndist,
                             float* out) {
    cvp::undistort_points(xy, n, fx, fy, cx, cy, dist, ndist, out);
}
void oracle_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int nPoints, int32_t* best) {
    for (int p = 0; p < nPoints; p++)
        best[p] = match_oracle::distinctive_descriptor(desc + 32 * (size_t)start[p], start[p + 1] - start[p]);
}

// ---- bag of words ---------------------------------------------------------------------------------
void* oracle_voc_create(int k, int L, int scoring, int weighting, int nNodes, const int32_t* parent,
                        const uint8_t* desc, const double* weight) {
    bow_oracle::Vocabulary* V = new bow_oracle::Vocabulary();
    V->build(k, L, scoring, weighting, nNodes, parent, desc, weight);
    return V;
}
void oracle_voc_destroy(void* h) { delete (bow_oracle::Vocabulary*)h; }
void oracle_bow_transform_features(void* h, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight,
                                   int32_t* nid) {
    const bow_oracle::Vocabulary& V = *(bow_oracle::Vocabulary*)h;
    for (int i = 0; i < n; i++) {
        int w, nd;
        bow_oracle::transform_feature(V, desc + 32 * (size_t)i, levelsup, w, weight[i], nd);
        word[i] = w; nid[i] = nd;
    }
}
// BowVector as (ids, values), FeatureVector as CSR (nodes, start, feat); caps: n entries each (+1 for start).
void oracle_bow_transform(void* h, const uint8_t* desc, int n, int levelsup, int32_t* nWords, uint32_t* ids,
                          double* values, int32_t* nNodes, uint32_t* nodes, int32_t* start, uint32_t* feat) {
    const bow_oracle::Vocabulary& V = *(bow_oracle::Vocabulary*)h;
    std::map<unsigned, double> bow;
    std::map<unsigned, std::vector<unsigned>> fv;
    bow_oracle::transform(V, desc, n, levelsup, bow, fv);
    int i = 0;
    for (auto& e : bow) { ids[i] = e.first; values[i] = e.second; i++; }
    *nWords = i;
    int j = 0, p = 0;
    for (auto& e : fv) {
        nodes[j] = e.first; start[j] = p;
        for (unsigned f : e.second) feat[p++] = f;
        j++;
    }
    start[j] = p;
    *nNodes = j;
}
static bow_oracle::FeatVec fill_fv(int nn, const int32_t* node, const int32_t* start, const int32_t* feat) {
    bow_oracle::FeatVec f;
    f.node.assign(node, node + nn);
    f.start.assign(start, start + nn + 1);
    f.feat.assign(feat, feat + start[nn]);
    return f;
}
int oracle_search_by_bow(int nnA, const int32_t* nodeA, const int32_t* startA, const int32_t* featA, const uint8_t* descA,
                         const float* angleA, const uint8_t* validA, int nA, int nnB, const int32_t* nodeB,
                         const int32_t* startB, const int32_t* featB, const uint8_t* descB, const float* angleB,
                         const uint8_t* validB, int nB, int thLow, int strict, float nnratio, int checkOri, int nLeftB,
                         int32_t* matchA, int32_t* matchAR) {
    const bow_oracle::FeatVec fa = fill_fv(nnA, nodeA, startA, featA), fb = fill_fv(nnB, nodeB, startB, featB);
    const bow_oracle::BowSearchParams prm{thLow, strict, nnratio, checkOri, nLeftB};
    return bow_oracle::search_by_bow(fa, descA, angleA, validA, nA, fb, descB, angleB, validB, nB, prm, matchA, matchAR);
}
int oracle_search_for_triangulation(int nnA, const int32_t* nodeA, const int32_t* startA, const int32_t* featA,
                                    const OrbKp* keysA, const uint8_t* descA, const float* urA, const uint8_t* mpA, int nA,
                                    int nnB, const int32_t* nodeB, const int32_t* startB, const int32_t* featB,
                                    const OrbKp* keysB, const uint8_t* descB, const float* urB, const uint8_t* mpB, int nB,
                                    const float* F12, const float* ep, const float* sfB, const float* sigma2B, int onlyStereo,
                                    int coarse, int checkOri, int thLow, int32_t* matches12) {
    const bow_oracle::FeatVec fa = fill_fv(nnA, nodeA, startA, featA), fb = fill_fv(nnB, nodeB, startB, featB);
    const bow_oracle::TriSide A{&fa, keysA, descA, urA, mpA, nA}, B{&fb, keysB, descB, urB, mpB, nB};
    return bow_oracle::search_for_triangulation(A, B, F12, ep, sfB, sigma2B, onlyStereo, coarse, checkOri, thLow, matches12);
}

// the two-camera form: rig = {nLeft1, nLeft2} + per combination {P1[8], P2[8], prec1, prec2, R12[9], t12[3]} x 4 as flat floats
int oracle_search_for_triangulation_rig(int nnA, const int32_t* nodeA, const int32_t* startA, const int32_t* featA,
                                        const OrbKp* keysA, const uint8_t* descA, const uint8_t* mpA, int nA, int nnB,
                                        const int32_t* nodeB, const int32_t* startB, const int32_t* featB, const OrbKp* keysB,
                                        const uint8_t* descB, const uint8_t* mpB, int nB, const float* sfB, const float* sigma2A,
                                        const float* sigma2B, int nLeft1, int nLeft2, const float* pairs /*[4][30]*/,
                                        int onlyStereo, int coarse, int checkOri, int thLow, int32_t* matches12) {
    const bow_oracle::FeatVec fa = fill_fv(nnA, nodeA, startA, featA), fb = fill_fv(nnB, nodeB, startB, featB);
    const bow_oracle::TriSide A{&fa, keysA, descA, nullptr, mpA, nA}, B{&fb, keysB, descB, nullptr, mpB, nB};
    bow_oracle::TriRig rig;
    rig.nLeft1 = nLeft1; rig.nLeft2 = nLeft2; rig.levelSigma2A = sigma2A;
    for (int k = 0; k < 4; k++) {
        const float* p = pairs + 30 * k;
        for (int i = 0; i < 8; i++) { rig.P1[k][i] = p[i]; rig.P2[k][i] = p[8 + i]; }
        rig.prec1[k] = p[16]; rig.prec2[k] = p[17];
        for (int i = 0; i < 9; i++) rig.R12[k][i] = p[18 + i];
        for (int i = 0; i < 3; i++) rig.t12[k][i] = p[27 + i];
    }
    const float F12[9] = {0}, ep[2] = {0, 0};
    return bow_oracle::search_for_triangulation(A, B, F12, ep, sfB, sigma2B, onlyStereo, coarse, checkOri, thLow, matches12, &rig);
}

}  // extern "C"
