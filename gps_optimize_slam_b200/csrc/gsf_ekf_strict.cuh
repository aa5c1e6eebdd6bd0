// Step-by-step EKF recursion for ONE trajectory, written once for host and device.
//
// This is the literal restatement of apply_ekf_correction (EKFGPSSLAM.py:831-935) with the
// diagonal-covariance specialisation (P0, Q, R are np.diag(...), F = I,
// H = [I3 0], so P stays diagonal and GPS updates never touch the quaternion):
//   relative pose   :77-92     predict :702-715     update :717-734
//   process_step    :736-772   outage state machine + RTS :875-928
//   sharp-turn gate :808-826
// It keeps the un-telescoped quaternion odometry (delta_q products, zero-motion fallback on
// a zero-norm quaternion), so it is the general path: the strict batched kernel runs it one
// thread per trajectory, and the CPU tests compile it with g++ to check it against the
// oracle without a GPU.  The RTS backward pass over an outage [s..i] is applied in closed
// form, x_s[k] = x_f[k] + P_f[k]/P_pred[i] * (x_f[i] - x_pred[i]), which is what the
// reference's recursion (:785-799) evaluates to when no update happens inside the outage.
#pragma once
#include "gsf_common.cuh"

namespace gsf {

// Sharp-turn gate over original SLAM quaternions [s..e] (EKFGPSSLAM.py:808-826).
GSF_HD inline bool sharp_turn_in_range(const double* ts, const double* quat, long s, long e, double thresh) {
    if (e - s + 1 < 2) return false;
    double worst = 0.0;
    for (long a = s + 1; a <= e; ++a) {
        double t1 = ts[a - 1], t2 = ts[a];
        if (t2 <= t1) continue;
        Quat q1{quat[4 * (a - 1)], quat[4 * (a - 1) + 1], quat[4 * (a - 1) + 2], quat[4 * (a - 1) + 3]};
        Quat q2{quat[4 * a], quat[4 * a + 1], quat[4 * a + 2], quat[4 * a + 3]};
        if (qnorm2(q1) == 0.0 || qnorm2(q2) == 0.0) return true;      // scipy ValueError -> True (:821)
        double y1 = yaw_zyx(q1), y2 = yaw_zyx(q2);
        double d = atan2(sin(y2 - y1), cos(y2 - y1));
        double rate = fabs(d / (t2 - t1));
        if (rate > worst) worst = rate;
    }
    return worst > thresh;
}

// Returns status bits.  out_pos[n,3], out_quat[n,4].
GSF_HD inline int ekf_strict_trajectory(long n, const double* ts, const double* pos, const double* quat,
                                        const double* z, const double* init_pos, const double* init_quat,
                                        const FuseParams& prm, double* out_pos, double* out_quat) {
    if (n <= 0) return ST_EMPTY;
    int status = ST_OK;
    double x[3] = {init_pos[0], init_pos[1], init_pos[2]};
    Quat qs = qunit_or_identity(Quat{init_quat[0], init_quat[1], init_quat[2], init_quat[3]});
    double P[3] = {prm.p0[0], prm.p0[1], prm.p0[2]};
    out_pos[0] = x[0]; out_pos[1] = x[1]; out_pos[2] = x[2];
    out_quat[0] = qs.x; out_quat[1] = qs.y; out_quat[2] = qs.z; out_quat[3] = qs.w;

    bool outage = row_has_nan(z[0], z[1], z[2]);
    long start = outage ? 0 : -1;
    double P_at_start[3] = {P[0], P[1], P[2]};        // P_f[start] (start == 0) or P_f[start-1]
    double t_last = ts[0];

    for (long i = 1; i < n; ++i) {
        double dt = fmax(1e-6, ts[i] - t_last);
        // ---- relative SLAM motion on the ORIGINAL poses (:77-92)
        Quat qa{quat[4 * (i - 1)], quat[4 * (i - 1) + 1], quat[4 * (i - 1) + 2], quat[4 * (i - 1) + 3]};
        Quat qb{quat[4 * i], quat[4 * i + 1], quat[4 * i + 2], quat[4 * i + 3]};
        double dpl[3] = {0.0, 0.0, 0.0};
        Quat dq{0.0, 0.0, 0.0, 1.0};
        if (qnorm2(qa) == 0.0 || qnorm2(qb) == 0.0) {
            status |= ST_BAD_QUATERNION;                                // zero motion (:84-86)
        } else {
            Quat q1 = qunit(qa), q2 = qunit(qb);
            double M1[9];
            qmat(q1, M1);
            matT_vec(M1, pos[3 * i] - pos[3 * (i - 1)], pos[3 * i + 1] - pos[3 * (i - 1) + 1],
                     pos[3 * i + 2] - pos[3 * (i - 1) + 2], dpl[0], dpl[1], dpl[2]);
            dq = qunit(qmul(qconj(q1), q2));
        }
        // ---- predict (:702-715)
        Quat qsu = qunit(qs);
        double Ms[9];
        qmat(qsu, Ms);
        double mx, my, mz;
        mat_vec(Ms, dpl[0], dpl[1], dpl[2], mx, my, mz);
        double xp[3] = {x[0] + mx, x[1] + my, x[2] + mz};
        Quat qp = qunit_or_identity(qunit(qmul(qsu, qunit(dq))));
        double dt_adj = fmax(fabs(dt), 1e-6);
        double Pp[3] = {P[0] + prm.q[0] * dt_adj, P[1] + prm.q[1] * dt_adj, P[2] + prm.q[2] * dt_adj};

        bool avail = !row_has_nan(z[3 * i], z[3 * i + 1], z[3 * i + 2]);
        bool do_rts = true;
        int steps_here = 0;
        if (!avail && !outage) {
            outage = true; start = i;
            P_at_start[0] = P[0]; P_at_start[1] = P[1]; P_at_start[2] = P[2];
        } else if (avail && outage) {
            if (sharp_turn_in_range(ts, quat, start, i - 1, prm.yaw_rate_thresh)) {
                do_rts = false;
                steps_here = prm.sharp_turn_steps;
            }
        }
        // ---- update / blend (:717-772)
        double xf[3] = {xp[0], xp[1], xp[2]};
        double Pf[3] = {Pp[0], Pp[1], Pp[2]};
        if (avail) {
            int eff = (avail && outage) ? steps_here : 0;
            double w = 1.0;
            if (eff > 0) { double wd = 1.0 / (double)eff; if (wd < 1.0) w = wd; }
            for (int a = 0; a < 3; ++a) {
                double k = Pp[a] * (1.0 / (Pp[a] + prm.r[a]));
                double xu = xp[a] + k * (z[3 * i + a] - xp[a]);
                double omk = 1.0 - k;
                Pf[a] = omk * Pp[a] * omk + k * prm.r[a] * k;
                xf[a] = (w < 1.0) ? ((1.0 - w) * xp[a] + w * xu) : xu;
            }
        }
        Quat qf = qp;                                   // K rows 3..6 are zero: q untouched (:728-729)
        out_pos[3 * i] = xf[0]; out_pos[3 * i + 1] = xf[1]; out_pos[3 * i + 2] = xf[2];
        out_quat[4 * i] = qf.x; out_quat[4 * i + 1] = qf.y; out_quat[4 * i + 2] = qf.z; out_quat[4 * i + 3] = qf.w;

        if (avail && outage) {
            if (do_rts && i - start + 1 > 1) {
                // closed-form RTS over [start .. i-1] (:785-799, :918-922)
                double Pk[3] = {P_at_start[0], P_at_start[1], P_at_start[2]};
                for (long k = start; k < i; ++k) {
                    if (k >= 1) {
                        double dk = fmax(fabs(fmax(1e-6, ts[k] - ts[k - 1])), 1e-6);
                        for (int a = 0; a < 3; ++a) Pk[a] += prm.q[a] * dk;
                    }
                    for (int a = 0; a < 3; ++a) out_pos[3 * k + a] += (Pk[a] / Pp[a]) * (xf[a] - xp[a]);
                }
            }
            outage = false; start = -1;
        }
        x[0] = xf[0]; x[1] = xf[1]; x[2] = xf[2];
        P[0] = Pf[0]; P[1] = Pf[1]; P[2] = Pf[2];
        qs = qf;
        t_last = ts[i];
    }
    return status;
}

}  // namespace gsf
