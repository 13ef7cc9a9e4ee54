// synth.cu -- seeded synthetic trajectories generated directly in HBM (bench / tests only).
//
// Same recipe as gochem_b200/synth.py: frame f = (reference + per-atom normal
// noise of sigma_a) rotated by a random unit quaternion and translated uniformly in
// [-50, 50]^3 A.  Randomness is a counter-based hash of (seed, frame id, atom, lane) so any frame
// can be generated independently on any GPU; it is NOT bit-identical to the numpy generator, so
// parity checks download the frames they compare.
#include "common.cuh"

namespace gcb {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double u01(unsigned long long h) {  // (0, 1)
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ unsigned long long key(unsigned long long seed, long frame, long atom, int lane) {
    unsigned long long k = mix64(seed ^ mix64((unsigned long long)frame * 0xD1B54A32D192ED03ull));
    k = mix64(k ^ ((unsigned long long)(atom + 7) * 0x8CB92BA72F3D8DD7ull));
    return mix64(k + (unsigned long long)lane);
}

__device__ __forceinline__ void normal2(unsigned long long k0, unsigned long long k1, double& g0, double& g1) {
    const double r = sqrt(-2.0 * log(u01(k0)));
    double s, c;
    sincospi(2.0 * u01(k1), &s, &c);
    g0 = r * c;
    g1 = r * s;
}

__global__ void __launch_bounds__(256) synth_kernel(double* __restrict__ dst, int natoms, int np, long nframes,
                                                   const double* __restrict__ ref_aos, const double* __restrict__ sigma,
                                                   unsigned long long seed, long frame_id0, int frame0_is_ref,
                                                   int through_f32) {
    const long total = nframes * (long)np;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        const long fl = g / np;
        const int i = (int)(g - fl * np);
        const long f = frame_id0 + fl;
        double x = 0.0, y = 0.0, z = 0.0;
        if (i < natoms) {
            const double bx = ref_aos[3L * i], by = ref_aos[3L * i + 1], bz = ref_aos[3L * i + 2];
            if (f == 0 && frame0_is_ref) {
                x = bx; y = by; z = bz;
            } else {
                double g0, g1, g2, g3;
                normal2(key(seed, f, i, 0), key(seed, f, i, 1), g0, g1);
                normal2(key(seed, f, i, 2), key(seed, f, i, 3), g2, g3);
                const double s = sigma[i];
                const double px = bx + s * g0, py = by + s * g1, pz = bz + s * g2;
                // per-frame rigid motion (same for every atom of the frame)
                double q0, q1, q2, q3;
                normal2(key(seed, f, -1, 0), key(seed, f, -1, 1), q0, q1);
                normal2(key(seed, f, -1, 2), key(seed, f, -1, 3), q2, q3);
                const double qn = 1.0 / sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
                const double w = q0 * qn, a = q1 * qn, b = q2 * qn, c = q3 * qn;
                const double tx = 100.0 * u01(key(seed, f, -2, 0)) - 50.0;
                const double ty = 100.0 * u01(key(seed, f, -2, 1)) - 50.0;
                const double tz = 100.0 * u01(key(seed, f, -2, 2)) - 50.0;
                // row vector times rotation matrix (rows as in synth.py quat_to_rot)
                const double m00 = 1 - 2 * (b * b + c * c), m01 = 2 * (a * b - w * c), m02 = 2 * (a * c + w * b);
                const double m10 = 2 * (a * b + w * c), m11 = 1 - 2 * (a * a + c * c), m12 = 2 * (b * c - w * a);
                const double m20 = 2 * (a * c - w * b), m21 = 2 * (b * c + w * a), m22 = 1 - 2 * (a * a + b * b);
                x = px * m00 + py * m10 + pz * m20 + tx;
                y = px * m01 + py * m11 + pz * m21 + ty;
                z = px * m02 + py * m12 + pz * m22 + tz;
            }
            if (through_f32) { x = (double)(float)x; y = (double)(float)y; z = (double)(float)z; }
        }
        double* d = dst + fl * 3L * np;
        d[i] = x;
        d[np + i] = y;
        d[2L * np + i] = z;
    }
}

int launch_synth(gcb_traj* t, long first, long nframes, const double* d_ref_aos, const double* d_sigma,
                 unsigned long long seed, long frame_id0, int frame0_is_ref, int through_f32) {
    if (nframes <= 0) return GCB_OK;
    long blocks = (nframes * (long)t->np + 255) / 256;
    const long cap = (long)t->sm_count * 16;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<(int)blocks, 256, 0, t->stream>>>(t->d_frames + first * 3L * t->np, t->natoms, t->np, nframes, d_ref_aos,
                                                     d_sigma, seed, frame_id0, frame0_is_ref, through_f32);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

}  // namespace gcb
