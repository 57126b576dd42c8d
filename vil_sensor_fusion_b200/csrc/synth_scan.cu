// libvlo_synth.so -- BENCH / TEST TOOL, not part of the hot path: synthesises spinning-lidar sweeps of an analytic
// scene directly in device memory, frame k from (seed, k), so that whole-bag runs (20 000 HDL-64 scan
// pairs, "generated on the fly per rank from (seed, frame id) so no input transfer") have no PCIe in the timed region.
// Same sensor model as vil_sensor_fusion_b200/synth.py (scene = inside of a box room + vertical cylinders + boxes;
// azimuth-major firing order, clockwise head, ROS sensor frame x forward / y left / z up), float32 ray casting; the
// trajectory is a closed loop so that any number of frames stays inside the room.  Rays without a return are NaN
// points (the organise kernel K0 drops non-finite points, as it does for a real driver's invalid returns).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

extern "C" {
typedef struct vlo_synth_scene {
    int n_planes, n_cyl, n_boxes;
    float planes[8][4];        // n.x n.y n.z d, n.p + d = 0, inward normals
    float cyl[16][3];          // cx cy r (infinite along z)
    float boxes[8][6];         // xmin ymin zmin xmax ymax zmax
} vlo_synth_scene;

typedef struct vlo_synth_sensor {
    int rings, n_az;           // HDL-64E: 64 x 1800
    float lower_deg, upper_deg;
    float max_range, noise_sigma, scan_period;
    int rolling;               // 1: every azimuth column from the pose at its own firing time
    float loop_a, loop_b, loop_period, yaw_wobble, bob_amp;     // trajectory: ellipse (a, b) once per loop_period seconds
} vlo_synth_sensor;
}

struct SynthPose { float R[9]; float p[3]; };

__device__ __forceinline__ SynthPose synth_pose(const vlo_synth_sensor &s, float t)
{
    const float w = 6.283185307179586f / s.loop_period, ph = w * t;
    float sp, cp;
    sincosf(ph, &sp, &cp);
    SynthPose o;
    o.p[0] = s.loop_a * cp; o.p[1] = s.loop_b * sp; o.p[2] = s.bob_amp * sinf(1.1f * t);
    const float yaw = atan2f(s.loop_b * cp, -s.loop_a * sp) + s.yaw_wobble * sinf(0.3f * t);     // heading = tangent + wobble
    float sy, cy;
    sincosf(yaw, &sy, &cy);
    o.R[0] = cy; o.R[1] = -sy; o.R[2] = 0.f; o.R[3] = sy; o.R[4] = cy; o.R[5] = 0.f; o.R[6] = 0.f; o.R[7] = 0.f; o.R[8] = 1.f;
    return o;
}

__device__ __forceinline__ uint32_t synth_hash(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256) synth_scan_kernel(vlo_synth_scene sc, vlo_synth_sensor s, int frame_first, unsigned seed, float4 *out)
{
    const int frame = frame_first + blockIdx.y;
    const int n = s.rings * s.n_az;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int col = i / s.rings, ring = i - col * s.rings;
    const float frac = (float)col / (float)s.n_az;
    const float t = frame * s.scan_period + (s.rolling ? frac * s.scan_period : 0.0f);
    const SynthPose P = synth_pose(s, t);
    const float el = (s.lower_deg + (s.upper_deg - s.lower_deg) * (float)ring / (float)(s.rings - 1)) * 0.017453292519943295f;
    const float az = -6.283185307179586f * frac;
    float se, ce, sa, ca;
    sincosf(el, &se, &ce); sincosf(az, &sa, &ca);
    const float d[3] = { ca * ce, sa * ce, se };                                   // sensor frame
    const float wd[3] = { P.R[0] * d[0] + P.R[1] * d[1] + P.R[2] * d[2], P.R[3] * d[0] + P.R[4] * d[1] + P.R[5] * d[2],
                          P.R[6] * d[0] + P.R[7] * d[1] + P.R[8] * d[2] };
    float best = INFINITY;
    for (int k = 0; k < sc.n_planes; k++) {
        const float den = wd[0] * sc.planes[k][0] + wd[1] * sc.planes[k][1] + wd[2] * sc.planes[k][2];
        const float num = -(P.p[0] * sc.planes[k][0] + P.p[1] * sc.planes[k][1] + P.p[2] * sc.planes[k][2] + sc.planes[k][3]);
        const float tt = num / den;
        if (den < 0.0f && tt > 1e-6f && tt < best) best = tt;
    }
    for (int k = 0; k < sc.n_cyl; k++) {
        const float ox = P.p[0] - sc.cyl[k][0], oy = P.p[1] - sc.cyl[k][1], r = sc.cyl[k][2];
        const float a = wd[0] * wd[0] + wd[1] * wd[1], b = 2.0f * (ox * wd[0] + oy * wd[1]), c = ox * ox + oy * oy - r * r;
        const float disc = b * b - 4.0f * a * c;
        if (disc > 0.0f && a > 1e-12f) {
            const float tt = (-b - sqrtf(disc)) / (2.0f * a);
            if (tt > 1e-6f && tt < best) best = tt;
        }
    }
    for (int k = 0; k < sc.n_boxes; k++) {
        float tmin = -INFINITY, tmax = INFINITY;
        #pragma unroll
        for (int a = 0; a < 3; a++) {
            const float inv = 1.0f / wd[a];
            const float t0 = (sc.boxes[k][a] - P.p[a]) * inv, t1 = (sc.boxes[k][3 + a] - P.p[a]) * inv;
            tmin = fmaxf(tmin, fminf(t0, t1)); tmax = fminf(tmax, fmaxf(t0, t1));
        }
        if (tmax >= tmin && tmin > 1e-6f && tmin < best) best = tmin;
    }
    float4 o = make_float4(NAN, NAN, NAN, 1.0f);
    if (best <= s.max_range) {
        if (s.noise_sigma > 0.0f) {
            const uint32_t h1 = synth_hash(seed, (uint32_t)frame, (uint32_t)i), h2 = synth_hash(seed ^ 0xA5A5A5A5u, (uint32_t)i, (uint32_t)frame);
            const float u1 = ((float)(h1 >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(h2 >> 8) * (1.0f / 16777216.0f);
            best += s.noise_sigma * sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
        }
        o = make_float4(d[0] * best, d[1] * best, d[2] * best, 1.0f);
    }
    out[(size_t)blockIdx.y * n + i] = o;
}

// pose of frame k (ROS frame): R row-major 9 + p 3 -> 12 floats per frame
__global__ void synth_pose_kernel(vlo_synth_sensor s, int frame_first, int n_frames, float *out12)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_frames) return;
    const SynthPose P = synth_pose(s, (frame_first + k) * s.scan_period);
    for (int a = 0; a < 9; a++) out12[k * 12 + a] = P.R[a];
    for (int a = 0; a < 3; a++) out12[k * 12 + 9 + a] = P.p[a];
}

extern "C" {

// Sweeps [frame_first, frame_first + n_frames) into d_out (device, n_frames x rings*n_az x 4 float32, x y z 1), enqueued on
// `stream` (a cudaStream_t, may be NULL).  Returns 0 or a cudaError_t.
int vlo_synth_scans(const vlo_synth_scene *scene, const vlo_synth_sensor *sensor, int frame_first, int n_frames, unsigned seed,
                    float *d_out, void *stream)
{
    if (!scene || !sensor || !d_out || n_frames < 1 || sensor->rings < 2 || sensor->n_az < 1 || scene->n_planes > 8 || scene->n_cyl > 16 ||
        scene->n_boxes > 8) return -1;
    const int n = sensor->rings * sensor->n_az;
    dim3 grid((n + 255) / 256, n_frames);
    synth_scan_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*scene, *sensor, frame_first, seed, (float4 *)d_out);
    return (int)cudaGetLastError();
}

// ground-truth poses (host out, 12 floats per frame); synchronous
int vlo_synth_poses(const vlo_synth_sensor *sensor, int frame_first, int n_frames, float *out12)
{
    if (!sensor || !out12 || n_frames < 1) return -1;
    float *d = nullptr;
    cudaError_t e = cudaMalloc((void **)&d, sizeof(float) * 12 * (size_t)n_frames);
    if (e != cudaSuccess) return (int)e;
    synth_pose_kernel<<<(n_frames + 127) / 128, 128>>>(*sensor, frame_first, n_frames, d);
    e = cudaMemcpy(out12, d, sizeof(float) * 12 * (size_t)n_frames, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return (int)e;
}

}
