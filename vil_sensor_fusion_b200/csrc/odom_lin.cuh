// Per-correspondence linearisation of the scan-to-scan step (LaserOdometry; oracle/laser_odometry.c
// is the frozen operation order): point-to-line / point-to-plane coefficients with upstream's robust weight, and the Jacobian
// row of the odometry step.  A header of its own so that the kernels (k3_odometry.cu) and the host emulation harness
// (tests/host/) compile the very same text.
#pragma once
#include "vlo_internal.cuh"

// Jacobian row of upstream's odometry step (s = 1), expression order as oracle/laser_odometry.c
__device__ __forceinline__ void odom_jacobian_row(const float *T, const float *trig, float x, float y, float z,
                                                  const float *coeff, float *row, float &bval)
{
    float srx = trig[0], crx = trig[1], sry = trig[2], cry = trig[3], srz = trig[4], crz = trig[5];
    float tx = T[3], ty = T[4], tz = T[5];
    float cx_ = coeff[0], cy_ = coeff[1], cz_ = coeff[2];
    float arx = (-crx * sry * srz * x + crx * crz * sry * y + srx * sry * z
                 + tx * crx * sry * srz - ty * crx * crz * sry - tz * srx * sry) * cx_
              + (srx * srz * x - crz * srx * y + crx * z
                 + ty * crz * srx - tz * crx - tx * srx * srz) * cy_
              + (crx * cry * srz * x - crx * cry * crz * y - cry * srx * z
                 + tz * cry * srx + ty * crx * cry * crz - tx * crx * cry * srz) * cz_;
    float ary = ((-crz * sry - cry * srx * srz) * x
                 + (cry * crz * srx - sry * srz) * y - crx * cry * z
                 + tx * (crz * sry + cry * srx * srz) + ty * (sry * srz - cry * crz * srx)
                 + tz * crx * cry) * cx_
              + ((cry * crz - srx * sry * srz) * x
                 + (cry * srz + crz * srx * sry) * y - crx * sry * z
                 + tz * crx * sry - ty * (cry * srz + crz * srx * sry)
                 - tx * (cry * crz - srx * sry * srz)) * cz_;
    float arz = ((-cry * srz - crz * srx * sry) * x + (cry * crz - srx * sry * srz) * y
                 + tx * (cry * srz + crz * srx * sry) - ty * (cry * crz - srx * sry * srz)) * cx_
              + (-crx * crz * x - crx * srz * y
                 + ty * crx * srz + tx * crx * crz) * cy_
              + ((cry * crz * srx - sry * srz) * x + (crz * sry + cry * srx * srz) * y
                 + tx * (sry * srz - cry * crz * srx) - ty * (crz * sry + cry * srx * srz)) * cz_;
    float atx = -(cry * crz - srx * sry * srz) * cx_ + crx * srz * cy_ - (crz * sry + cry * srx * srz) * cz_;
    float aty = -(cry * srz + crz * srx * sry) * cx_ - crx * crz * cy_ - (sry * srz - cry * crz * srx) * cz_;
    float atz = crx * sry * cx_ - srx * cy_ - crx * cry * cz_;
    row[0] = arx; row[1] = ary; row[2] = arz; row[3] = atx; row[4] = aty; row[5] = atz;
    bval = (float)(-0.05 * (double)coeff[3]);
}

__device__ __forceinline__ bool edge_coeff(float4 sel, float4 a, float4 b, int iter, float *coeff)
{
    float x0 = sel.x, y0 = sel.y, z0 = sel.z, x1 = a.x, y1 = a.y, z1 = a.z, x2 = b.x, y2 = b.y, z2 = b.z;
    float m1 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    float m2 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    float m3 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    float a012 = sqrtf(m1 * m1 + m2 * m2 + m3 * m3);
    float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    float la = ((y1 - y2) * m1 + (z1 - z2) * m2) / a012 / l12;
    float lb = -((x1 - x2) * m1 - (z1 - z2) * m3) / a012 / l12;
    float lc = -((x1 - x2) * m2 + (y1 - y2) * m3) / a012 / l12;
    float ld2 = a012 / l12;
    float s = 1.0f;
    if (iter >= 5) s = 1.0f - 1.8f * fabsf(ld2);
    coeff[0] = s * la; coeff[1] = s * lb; coeff[2] = s * lc; coeff[3] = s * ld2;
    return (double)s > 0.1 && ld2 != 0.0f;
}

__device__ __forceinline__ bool plane_coeff(float4 sel, float4 t1, float4 t2, float4 t3, int iter, float *coeff)
{
    float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
    float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
    float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
    float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa = pa / ps; pb = pb / ps; pc = pc / ps; pd = pd / ps;
    float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
    float s = 1.0f;
    if (iter >= 5) {
        float dist = sqrtf(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z);
        s = 1.0f - 1.8f * fabsf(pd2) / sqrtf(dist);
    }
    coeff[0] = s * pa; coeff[1] = s * pb; coeff[2] = s * pc; coeff[3] = s * pd2;
    return (double)s > 0.1 && pd2 != 0.0f;
}
