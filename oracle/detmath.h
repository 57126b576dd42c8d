/* TEST INFRASTRUCTURE -- CPU oracle, never linked into the product (libvlo.so).
 *
 * Deterministic float32 elementary functions used wherever a transcendental feeds a
 * quantity that must be bit-exact between this oracle and the CUDA path (ring ids,
 * rel-time, per-point de-skew rotations, pose trig).  The reference (LOAM fork, see
 * SURVEY.md Appendix A.1/A.4) calls libm's std::atan / std::atan2 / sin / cos; libm and
 * CUDA's libdevice do not agree to the last ulp, so the algorithm is FROZEN here as the
 * classic Cephes single-precision kernels (Moshier, cephes/single: sinf.c, atanf.c),
 * evaluated with separate IEEE mul/add (compile with -ffp-contract=off; the CUDA side
 * is compiled with --fmad=false).  Max error ~2 ulp, same class as libm.
 */
#ifndef ORC_DETMATH_H
#define ORC_DETMATH_H
#include <math.h>

static inline void orc_sincosf(float x, float *s_out, float *c_out)
{
    /* Cody-Waite reduction by pi/2 (three-term constant), then degree-7/8 minimax kernels. */
    const float two_over_pi = 0.63661977236758134308f;
    const float P1 = 1.5703125f;               /* pi/2 split: exact in 8 bits   */
    const float P2 = 4.837512969970703125e-4f;
    const float P3 = 7.54978995489188216e-8f;
    float kf = rintf(x * two_over_pi);
    int k = (int)kf;
    float r = x - kf * P1;
    r = r - kf * P2;
    r = r - kf * P3;
    float z = r * r;
    float sp = -1.9515295891e-4f * z;
    sp = sp + 8.3321608736e-3f;
    sp = sp * z;
    sp = sp - 1.6666654611e-1f;
    sp = sp * z;
    sp = sp * r;
    float sn = sp + r;
    float cp = 2.443315711809948e-5f * z;
    cp = cp - 1.388731625493765e-3f;
    cp = cp * z;
    cp = cp + 4.166664568298827e-2f;
    cp = cp * z;
    cp = cp * z;
    float hz = 0.5f * z;
    float cs = cp - hz;
    cs = cs + 1.0f;
    switch (k & 3) {
    case 0: *s_out = sn;  *c_out = cs;  break;
    case 1: *s_out = cs;  *c_out = -sn; break;
    case 2: *s_out = -sn; *c_out = -cs; break;
    default:*s_out = -cs; *c_out = sn;  break;
    }
}

static inline float orc_atanf(float xx)
{
    float x = fabsf(xx), y;
    if (x > 2.414213562373095f) {          /* tan(3pi/8) */
        y = 1.5707963267948966f;
        x = -(1.0f / x);
    } else if (x > 0.4142135623730950f) {  /* tan(pi/8) */
        y = 0.7853981633974483f;
        x = (x - 1.0f) / (x + 1.0f);
    } else {
        y = 0.0f;
    }
    float z = x * x;
    float p = 8.05374449538e-2f * z;
    p = p - 1.38776856032e-1f;
    p = p * z;
    p = p + 1.99777106478e-1f;
    p = p * z;
    p = p - 3.33329491539e-1f;
    p = p * z;
    p = p * x;
    p = p + x;
    y = y + p;
    return (xx < 0.0f) ? -y : y;
}

static inline float orc_atan2f(float y, float x)
{
    const float PI_F = 3.14159265358979323846f;
    const float PIO2_F = 1.5707963267948966f;
    if (x == 0.0f) {
        if (y > 0.0f) return PIO2_F;
        if (y < 0.0f) return -PIO2_F;
        return 0.0f;
    }
    float a = orc_atanf(y / x);
    if (x < 0.0f) {
        if (y < 0.0f) return a - PI_F;
        return a + PI_F;
    }
    return a;
}
#endif
