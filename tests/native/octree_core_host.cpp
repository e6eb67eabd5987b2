// tests/native/octree_core_host.cpp -- compiles the product's CTA-parallel octree
// (orb-slam3_byzyh_b200/csrc/octree_core.h) for the HOST as a single "thread" so its logic
// (list order, phase-2 std::sort emulation) can be unit-tested on CPU against the oracle.
// This is a test of product logic, not a CPU path of the product.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../orb-slam3_byzyh_b200/csrc/octree_core.h"

extern "C" int octree_core_host(const int* xys, int n, int minX, int maxX, int minY, int maxY, int N,
                                int* keep, int cap) {
    const int width = maxX - minX, height = maxY - minY;
    const int nIni = (int)roundf((float)width / height);
    const float hX = (float)width / nIni;
    int M = (N + 3 > 4 * nIni ? N + 3 : 4 * nIni) + 1;
    std::vector<uint32_t> pk(n), pnode(n);
    for (int i = 0; i < n; i++) pk[i] = OC_PACK(xys[3 * i], xys[3 * i + 1], xys[3 * i + 2]);
    std::vector<char> mem(oc_shared_bytes(M));
    OcWork w;
    oc_carve(w, mem.data(), M);
    w.pk = pk.data(); w.pnode = pnode.data(); w.n = n;
    int outn = 0;
    oc_distribute(w, width, height, nIni, hX, N, w.cc, &outn, w.cpos);   // same aliasing as the kernel
    for (int i = 0; i < outn && i < cap; i++) keep[i] = w.cc[i];
    return outn;
}

// std::sort emulation alone: sorts key<<32|payload words.
extern "C" void octree_core_sort(uint64_t* a, int n) { oc_std_sort(a, n); }

// Reference behaviour for the emulation tests: the real libstdc++ algorithms.
#include <algorithm>
static bool key_less(const uint64_t& a, const uint64_t& b) { return (a >> 32) < (b >> 32); }
extern "C" void octree_core_sort_ref(uint64_t* a, int n) { std::sort(a, a + n, key_less); }
extern "C" void octree_core_heapsort(uint64_t* a, int n) { oc_heap_sort(a, n); }
extern "C" void octree_core_heapsort_ref(uint64_t* a, int n) { std::partial_sort(a, a + n, a + n, key_less); }

// The CTA form of the same sort (level-parallel introsort loop + rank pass), here with one "thread".
extern "C" void octree_core_sort_cta(uint64_t* a, uint64_t* out, int n) {
    std::vector<int> q(4 * (n / 16 + 2));
    int cnt[3];
    oc_std_sort_cta(a, out, n, q.data(), cnt);
}
