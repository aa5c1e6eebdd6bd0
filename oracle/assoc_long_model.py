"""Test infrastructure (never imported by the product): numpy model of the LOCAL not-a-knot spline solve of
csrc/gsf_assoc_long.cu -- every chunk of CH knots eliminates forward from H knots before it, eliminates from H knots
behind it down to the knot after the chunk, joins the two at its last row (a 2 x 2 system) and back-substitutes; a natural
end (m = 0) where the halo cuts the segment, the true not-a-knot rows where the segment really ends; only INTERIOR moments
are produced (the end moments follow from the not-a-knot rows).  tests/test_kernel_math_cpu.py compares it with the global
solve (scipy CubicSpline, the spline of interp1d(kind='cubic'), EKFGPSSLAM.py:351-380): the halo bound 0.268^H."""
import numpy as np


def local_halo_moments(t, y, gap=5.0, CH=13, H=20):
    """t [M] increasing, y [M] -> m [M] second derivatives at the interior knots of every segment with >= 4 knots (NaN at
    segment ends and in shorter segments)."""
    M = len(t)
    m = np.full(M, np.nan)
    h = np.diff(t)
    seg_start = np.concatenate([[0], np.nonzero(h > gap)[0] + 1])
    seg_end = np.concatenate([seg_start[1:] - 1, [M - 1]])
    for j0 in range(0, M, CH):
        j1 = min(j0 + CH, M) - 1
        for sa, sb in zip(seg_start, seg_end):
            if sb < j0 or sa > j1 or sb - sa + 1 < 4:
                continue
            a, b = max(sa, j0 - H), min(sb, j1 + H)
            lt, rt = a == sa, b == sb
            U0, U1 = a + 1, b - 1
            R0, R1 = max(max(a, j0), U0), min(min(b, j1), U1)
            if R0 > R1:
                continue

            def row(j):
                lo, di, up = h[j - 1], 2.0 * (h[j - 1] + h[j]), h[j]
                if j == U0:
                    if lt:
                        di += h[a] * (h[a] + h[a + 1]) / h[a + 1]; up -= h[a] * h[a] / h[a + 1]
                    lo = 0.0
                if j == U1:
                    if rt:
                        di += h[b - 1] * (h[b - 2] + h[b - 1]) / h[b - 2]; lo -= h[b - 1] * h[b - 1] / h[b - 2]
                    up = 0.0
                rhs = 6.0 * ((y[j + 1] - y[j]) / h[j] - (y[j] - y[j - 1]) / h[j - 1])
                return lo, di, up, rhs

            c, d = {}, {}
            pc = pd = 0.0
            for j in range(U0, R1 + 1):                         # forward elimination
                lo, di, up, rhs = row(j)
                den = di - lo * pc
                pc, pd = up / den, (rhs - lo * pd) / den
                c[j], d[j] = pc, pd
            bp = e = 0.0
            for j in range(U1, R1, -1):                         # elimination from the far end
                lo, di, up, rhs = row(j)
                den = di - up * bp
                bp, e = lo / den, (rhs - up * e) / den
            mj = (d[R1] - c[R1] * e) / (1.0 - c[R1] * bp) if U1 > R1 else d[R1]
            m[R1] = mj
            for j in range(R1 - 1, R0 - 1, -1):                 # back-substitution through the chunk
                mj = d[j] - c[j] * mj
                m[j] = mj
    return m
