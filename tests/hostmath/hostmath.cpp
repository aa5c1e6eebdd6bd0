// Host build of the kernels' scalar math (gsf_common.cuh / gsf_ekf_strict.cuh compiled with
// g++), exported with a C ABI so the CPU test-suite can compare it with numpy/scipy and the
// oracle when no GPU is present.  Test infrastructure only.
#include "../../gps_optimize_slam_b200/csrc/gsf_common.cuh"
#include "../../gps_optimize_slam_b200/csrc/gsf_ekf_strict.cuh"

extern "C" {
int hm_umeyama_finish(int n, const double* mu_s, const double* mu_d, const double* H, double ss,
                      double* R, double* t, double* s) {
    return gsf::umeyama_finish(n, mu_s, mu_d, H, ss, R, t, *s);
}
void hm_quat_from_matrix(const double* m, double* q) {
    gsf::Quat r = gsf::quat_from_matrix(m);
    q[0] = r.x; q[1] = r.y; q[2] = r.z; q[3] = r.w;
}
double hm_yaw_zyx(const double* q) { return gsf::yaw_zyx(gsf::Quat{q[0], q[1], q[2], q[3]}); }
int hm_ekf_strict(long n, const double* ts, const double* pos, const double* quat, const double* z,
                  const double* init_pos, const double* init_quat, const double* params,
                  double* out_pos, double* out_quat) {
    const gsf::FuseParams* prm = reinterpret_cast<const gsf::FuseParams*>(params);
    return gsf::ekf_strict_trajectory(n, ts, pos, quat, z, init_pos, init_quat, *prm, out_pos, out_quat);
}
}

#include "../../gps_optimize_slam_b200/csrc/gsf_text.cuh"
extern "C" {
int hm_parse_double(const char* s, int n, double* out, int* inexact) { return gsf::parse_double(s, s + n, out, inexact); }
int hm_format_fixed(double x, int D, char* buf, int* bad) { return gsf::format_fixed(x, D, buf, bad); }
}
