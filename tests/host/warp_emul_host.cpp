// TEST INFRASTRUCTURE -- lock-step emulation of ONE WARP on the CPU: 32 host threads, __syncwarp = a barrier.
// Enough to run the warp-level device functions of csrc/dense6.cuh that only use __syncwarp and the lane index
// (single-warp 6x6 Jacobi, the Gauss-Newton update with the degeneracy projection) exactly as the GPU schedules them, and
// to compare them with the oracle's sequential C (tests/test_host_grid_search.py).
#include <cuda_runtime.h>
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

static std::barrier<> *g_bar = nullptr;
static inline void __syncwarp(unsigned = 0xffffffffu) { g_bar->arrive_and_wait(); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
using std::isfinite;
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

#include "../../vil_sensor_fusion_b200/csrc/dense6.cuh"

template <class F> static void run_warp(F f)
{
    std::barrier<> bar(32);
    g_bar = &bar;
    std::vector<std::thread> th;
    for (int lane = 0; lane < 32; lane++) th.emplace_back([&, lane] { f(lane); });
    for (auto &t : th) t.join();
    g_bar = nullptr;
}

extern "C" {
// eigenvalues ascending, evec row i = eigenvector i (vlo_eig6_jacobi_warp)
void host_eig6_warp(const float *A36, float *eval6, float *evec36)
{
    static GnScratch S;
    memcpy(S.A, A36, sizeof(float) * 36);
    run_warp([&](int lane) { vlo_eig6_jacobi_warp(S.A, S.V, S.cs_sn, S.eval, S.evec, lane); });
    memcpy(eval6, S.eval, sizeof(float) * 6); memcpy(evec36, S.evec, sizeof(float) * 36);
}
// one Gauss-Newton update from the 28 reduced totals (vlo_gn_update_warp); T, P, is_degenerate carry over between iterations
void host_gn_update_warp(const float *total28, int iter, float degen_thr, float dT_abort, float dR_abort, float *T6, float *P36,
                         int *is_degenerate, float *eval6, int *converged)
{
    static GnScratch S;
    memcpy(S.total, total28, sizeof(float) * VLO_NTERM); memcpy(S.T, T6, sizeof(float) * 6); memcpy(S.P, P36, sizeof(float) * 36);
    S.is_degenerate = *is_degenerate; S.converged = 0;
    run_warp([&](int lane) { vlo_gn_update_warp(S, iter, degen_thr, dT_abort, dR_abort, lane); });
    memcpy(T6, S.T, sizeof(float) * 6); memcpy(P36, S.P, sizeof(float) * 36);
    if (iter == 0) memcpy(eval6, S.eval, sizeof(float) * 6);
    *is_degenerate = S.is_degenerate; *converged = S.converged;
}
}
