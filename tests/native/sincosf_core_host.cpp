// Host build of csrc/sincosf_core.h for tests/test_sincosf_core.py.
#include "../../orb-slam3_byzyh_b200/csrc/sincosf_core.h"
extern "C" void host_sincosf(const float* y, int n, float* sn, float* cs) {
    for (int i = 0; i < n; i++) glibc_sincosf(y[i], sn + i, cs + i);
}
