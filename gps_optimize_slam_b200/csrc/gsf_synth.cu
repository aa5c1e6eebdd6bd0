// Synthetic trajectory generator (device side), for the large bench configs that cannot be
// produced on the host (config 3: 2^20 trajectories x 1000 poses = 92 GB of inputs).
// Same model as gps_optimize_slam_b200/synth.py: planar random-walk-yaw vehicle
// in the world/UTM frame, SLAM copy in a Sim3-related frame with integrated odometry drift,
// GNSS = world + white noise, pre-associated (GNSS stamps = SLAM stamps).  Randomness is a
// counter-based hash (splitmix64 of (seed, trajectory, step, stream)), so any slice of the
// batch can be regenerated independently and a slice copied to the host feeds the oracle.
#include "gsf_common.cuh"
#include "gsf_internal.cuh"

namespace gsf {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
struct Rng {
    uint64_t key; uint64_t ctr;
    __device__ Rng(uint64_t seed, uint64_t traj) : key(mix64(seed ^ mix64(traj * 0x632be59bd9b4e019ull + 1))), ctr(0) {}
    __device__ double uniform() {                       // (0,1)
        uint64_t r = mix64(key + (ctr++) * 0xd1342543de82ef95ull);
        return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    }
    __device__ void normal2(double& a, double& b) {     // Box-Muller
        double u1 = uniform(), u2 = uniform();
        double r = sqrt(-2.0 * log(u1)), s, c;
        sincospi(2.0 * u2, &s, &c);
        a = r * c; b = r * s;
    }
};


__global__ void synth_kernel(const SynthArgs A) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= A.B) return;
    const double PI = 3.141592653589793238462643383279502884;
    Rng rng(A.seed, (uint64_t)(A.traj0 + b));
    const long long e0 = (long long)b * A.n;
    double yaw = (2.0 * rng.uniform() - 1.0) * PI;
    double wx = 455779.0 + (2.0 * rng.uniform() - 1.0) * 1e3;
    double wy = 5431368.0 + (2.0 * rng.uniform() - 1.0) * 1e3;
    double wz = 112.0 + (2.0 * rng.uniform() - 1.0) * 10.0;
    const double s_gt = 0.9 + 0.2 * rng.uniform();
    const double a_gt = (2.0 * rng.uniform() - 1.0) * PI;
    const double sigma_g = 0.05 + 0.45 * rng.uniform();
    // small fixed tilt so that the quaternions are not purely about z
    double tx, ty; rng.normal2(tx, ty);
    Quat tilt = qunit(Quat{0.02 * tx, 0.02 * ty, 0.0, 1.0});
    int out_a = -1, out_b = -1;
    if (rng.uniform() < A.outage_prob && A.n > 8) {
        int len = 1 + (int)(rng.uniform() * A.outage_max_len);
        out_a = (int)(rng.uniform() * (A.n - 2));
        out_b = min(A.n, out_a + len);
    }
    const double ca = cos(a_gt), sa = sin(a_gt);
    const double t0x = wx, t0y = wy, t0z = wz;
    double yaw_rate = 0.0, dpx = 0.0, dpy = 0.0, dpz = 0.0, dyaw = 0.0;
    for (int i = 0; i < A.n; ++i) {
        double n0, n1, n2, n3, n4, n5, n6, n7;
        rng.normal2(n0, n1); rng.normal2(n2, n3); rng.normal2(n4, n5); rng.normal2(n6, n7);
        yaw_rate += 0.004 * n0;
        yaw += yaw_rate * A.dt;
        if (i > 0) {
            wx += A.speed * cos(yaw) * A.dt; wy += A.speed * sin(yaw) * A.dt;
            wz += A.speed * 0.02 * sin(0.05 * i) * A.dt;
        }
        dpx += 0.02 * n1; dpy += 0.02 * n2; dpz += 0.004 * n3;
        dyaw += (0.2 * PI / 180.0) * n4;
        // slam_true = Rg^T (w - t) / s
        const double rx = (wx - t0x) / s_gt, ry = (wy - t0y) / s_gt, rz = (wz - t0z) / s_gt;
        const long long e = e0 + i;
        A.ts[e] = i * A.dt;
        A.pos[3 * e] = ca * rx + sa * ry + dpx;
        A.pos[3 * e + 1] = -sa * rx + ca * ry + dpy;
        A.pos[3 * e + 2] = rz + dpz;
        const double th = 0.5 * (yaw - a_gt + dyaw);
        Quat q = qmul(Quat{0.0, 0.0, sin(th), cos(th)}, tilt);
        A.quat[4 * e] = q.x; A.quat[4 * e + 1] = q.y; A.quat[4 * e + 2] = q.z; A.quat[4 * e + 3] = q.w;
        const bool out = (i >= out_a && i < out_b);
        A.z[3 * e] = out ? nan("") : wx + sigma_g * n5;
        A.z[3 * e + 1] = out ? nan("") : wy + sigma_g * n6;
        A.z[3 * e + 2] = out ? nan("") : wz + sigma_g * n7;
    }
}

cudaError_t launch_synth(const SynthArgs& a, cudaStream_t stream) {
    if (a.B <= 0) return cudaSuccess;
    synth_kernel<<<(a.B + 127) / 128, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace gsf
