// N4: the on-disk formats on the device.
//   gsf_parse_table_dev   np.loadtxt of a numeric table (/root/reference/EKFGPSSLAM.py:110-125 load_slam_trajectory, :252-258
//                         load_gps_data): lines split at '\n', '#' starts a comment, blank lines are skipped, fields are
//                         separated by runs of whitespace (delimiter 0) or by exactly one delimiter character (' ' or ','
//                         as the reference tries them, :252-253); every field goes through parse_double (gsf_text.cuh:
//                         strtod-quality, so the doubles equal numpy's bit for bit).
//   gsf_write_pose_rows_dev  the two np.savetxt calls of :1087-1102: rows "ts a b c qx qy qz qw" with per-column "%.Df".
// Both are two-pass (count / exclusive scan / emit): newline positions -> line starts -> valid rows -> parsed rows, and row
// lengths -> byte offsets -> text.  HBM-bound byte work; one thread per line / row.
#include "gsf_text.cuh"
#include "gsf_internal.cuh"

namespace gsf {

constexpr int TX_T = 256;
constexpr int TX_BYTES = 64;                       // bytes per thread in the newline passes
constexpr int TX_CHUNK = TX_T * TX_BYTES;

// exclusive scan of `n` 64-bit counts by ONE block (n = number of first-level blocks); total -> *total
__global__ void __launch_bounds__(1024) scan_counts_kernel(long long* __restrict__ v, long long n, long long* __restrict__ total) {
    __shared__ long long wsum[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long x = i < n ? v[i] : 0;
        long long inc = x;
        for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(GSF_FULL_MASK, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        long long pre = carry;
        for (int w = 0; w < warp; ++w) pre += wsum[w];
        if (i < n) v[i] = pre + inc - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry = pre + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

template <typename F>
__device__ __forceinline__ long long block_exclusive_count(long long mine, F&& unused, long long* wsum_smem) {
    (void)unused;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long inc = mine;
    for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(GSF_FULL_MASK, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) wsum_smem[warp] = inc;
    __syncthreads();
    long long pre = 0;
    for (int w = 0; w < warp; ++w) pre += wsum_smem[w];
    __syncthreads();
    return pre + inc - mine;
}

// ---- parser pass 1: newlines per chunk
__global__ void __launch_bounds__(TX_T) text_count_newlines_kernel(const char* __restrict__ text, long long nbytes, long long* __restrict__ counts) {
    __shared__ int tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    const long long b0 = (long long)blockIdx.x * TX_CHUNK + (long long)threadIdx.x * TX_BYTES;
    int c = 0;
    for (int k = 0; k < TX_BYTES; ++k) { const long long i = b0 + k; if (i < nbytes && text[i] == '\n') ++c; }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(GSF_FULL_MASK, c, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&tot, c);
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = tot;
}
// ---- parser pass 2: start offset of every line (line 0 starts at 0, line k after the k-th newline)
__global__ void __launch_bounds__(TX_T) text_line_starts_kernel(const char* __restrict__ text, long long nbytes, const long long* __restrict__ counts,
                                                                long long* __restrict__ line_start) {
    __shared__ long long wsum[TX_T / 32];
    const long long b0 = (long long)blockIdx.x * TX_CHUNK + (long long)threadIdx.x * TX_BYTES;
    int c = 0;
    for (int k = 0; k < TX_BYTES; ++k) { const long long i = b0 + k; if (i < nbytes && text[i] == '\n') ++c; }
    long long line = counts[blockIdx.x] + block_exclusive_count((long long)c, 0, wsum) + 1;      // index of the next line to start
    if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
    for (int k = 0; k < TX_BYTES; ++k) { const long long i = b0 + k; if (i < nbytes && text[i] == '\n') line_start[line++] = i + 1; }
}

__device__ __forceinline__ bool tx_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// ---- parser pass 3: valid (non-blank, non-comment) lines per block of TX_T lines
__global__ void __launch_bounds__(TX_T) text_classify_kernel(const char* __restrict__ text, long long nbytes, const long long* __restrict__ line_start,
                                                             long long nlines, unsigned char* __restrict__ valid, long long* __restrict__ counts) {
    __shared__ int tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    const long long l = (long long)blockIdx.x * TX_T + threadIdx.x;
    int v = 0;
    if (l < nlines) {
        long long p = line_start[l];
        const long long e = l + 1 < nlines ? line_start[l + 1] - 1 : nbytes;
        while (p < e && tx_space(text[p])) ++p;
        v = (p < e && text[p] != '#' && text[p] != '\n') ? 1 : 0;
        valid[l] = (unsigned char)v;
    }
    const unsigned bal = __ballot_sync(GSF_FULL_MASK, v);
    if ((threadIdx.x & 31) == 0) atomicAdd(&tot, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = tot;
}
// ---- parser pass 4: parse every valid line into its row
// info: [0] rows, [1] min columns, [2] max columns, [3] status bits (1 unparsable field, 2 more than max_cols columns, 4 digits dropped
//       beyond 19 significant with an ambiguous rounding, 8 more rows than max_rows)
__global__ void __launch_bounds__(TX_T) text_parse_rows_kernel(const char* __restrict__ text, long long nbytes, const long long* __restrict__ line_start,
                                                               long long nlines, const unsigned char* __restrict__ valid,
                                                               const long long* __restrict__ counts, int delim, int max_cols, long long max_rows,
                                                               double* __restrict__ out, long long* __restrict__ info) {
    __shared__ long long wsum[TX_T / 32];
    const long long l = (long long)blockIdx.x * TX_T + threadIdx.x;
    const int v = l < nlines ? valid[l] : 0;
    const long long row = counts[blockIdx.x] + block_exclusive_count((long long)v, 0, wsum);
    if (!v) return;
    if (row >= max_rows) { atomicOr((unsigned long long*)(info + 3), 8ull); return; }
    const char* p = text + line_start[l];
    const char* e = text + (l + 1 < nlines ? line_start[l + 1] - 1 : nbytes);
    for (const char* c = p; c < e; ++c) if (*c == '#') { e = c; break; }          // comment: rest of the line
    while (e > p && tx_space(e[-1])) --e;
    while (p < e && tx_space(*p)) ++p;
    int ncol = 0, st = 0;
    double* o = out + row * (long long)max_cols;
    while (p < e) {
        const char* fe = p;                                                       // field end
        if (delim == 0) { while (fe < e && !tx_space(*fe)) ++fe; }
        else { while (fe < e && *fe != (char)delim) ++fe; }
        const char* a = p; const char* b = fe;
        while (a < b && tx_space(*a)) ++a;
        while (b > a && tx_space(b[-1])) --b;
        double val = 0.0; int inexact = 0;
        const int used = parse_double(a, b, &val, &inexact);
        if (used == 0 || a + used != b) st |= 1;                                  // float('') / float('abc'): ValueError in numpy
        if (inexact) st |= 4;
        if (ncol < max_cols) o[ncol] = val; else st |= 2;
        ++ncol;
        p = fe;
        if (delim == 0) { while (p < e && tx_space(*p)) ++p; }
        else if (p < e) { ++p; if (p == e) { st |= 1; ++ncol; } }                // trailing delimiter: an empty last field
    }
    atomicMin((unsigned long long*)(info + 1), (unsigned long long)ncol);
    atomicMax((unsigned long long*)(info + 2), (unsigned long long)ncol);
    if (st) atomicOr((unsigned long long*)(info + 3), (unsigned long long)st);
}
__global__ void text_info_init_kernel(long long* info, const long long* total_rows) {
    info[0] = total_rows ? *total_rows : 0; info[1] = 0x7fffffffffffffffll; info[2] = 0; info[3] = 0;
}

static long long div_up(long long a, long long b) { return (a + b - 1) / b; }

// work layout (bytes): counts1[nb1] | total1 | line_start[nlines_max] | valid[nlines_max] | counts2[nb2] | total2
long long parse_table_work_bytes(long long nbytes) {
    const long long nb1 = div_up(nbytes > 0 ? nbytes : 1, TX_CHUNK);
    const long long lines_max = nbytes + 1;                                       // every byte a newline
    const long long nb2 = div_up(lines_max, TX_T);
    return 8 * (nb1 + 2) + 8 * (lines_max + 1) + ((lines_max + 7) & ~7ll) + 8 * (nb2 + 2) + 64;
}
// Two host synchronisations inside (the line count sizes the later launches): this entry point is a loader, not a stream op.
cudaError_t launch_parse_table(const char* text, long long nbytes, int delim, int max_cols, double* out, long long max_rows, void* work,
                               long long* info, cudaStream_t stream) {
    const long long nb1 = div_up(nbytes > 0 ? nbytes : 1, TX_CHUNK);
    const long long lines_max = nbytes + 1;
    char* w = reinterpret_cast<char*>(work);
    long long* counts1 = reinterpret_cast<long long*>(w); w += 8 * (nb1 + 1);
    long long* total1 = reinterpret_cast<long long*>(w); w += 8;
    long long* line_start = reinterpret_cast<long long*>(w); w += 8 * (lines_max + 1);
    unsigned char* valid = reinterpret_cast<unsigned char*>(w); w += (lines_max + 7) & ~7ll;
    long long* counts2 = reinterpret_cast<long long*>(w);
    if (nbytes <= 0) { text_info_init_kernel<<<1, 1, 0, stream>>>(info, nullptr); return cudaGetLastError(); }
    text_count_newlines_kernel<<<(unsigned)nb1, TX_T, 0, stream>>>(text, nbytes, counts1);
    scan_counts_kernel<<<1, 1024, 0, stream>>>(counts1, nb1, total1);
    text_line_starts_kernel<<<(unsigned)nb1, TX_T, 0, stream>>>(text, nbytes, counts1, line_start);
    long long newlines = 0;
    cudaError_t e = cudaMemcpyAsync(&newlines, total1, 8, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    const long long nlines = newlines + 1;
    const long long nb2 = div_up(nlines, TX_T);
    long long* total2 = counts2 + nb2 + 1;
    text_classify_kernel<<<(unsigned)nb2, TX_T, 0, stream>>>(text, nbytes, line_start, nlines, valid, counts2);
    scan_counts_kernel<<<1, 1024, 0, stream>>>(counts2, nb2, total2);
    text_info_init_kernel<<<1, 1, 0, stream>>>(info, total2);
    text_parse_rows_kernel<<<(unsigned)nb2, TX_T, 0, stream>>>(text, nbytes, line_start, nlines, valid, counts2, delim, max_cols, max_rows, out, info);
    return cudaGetLastError();
}

// ---- writer: rows "ts a b c qx qy qz qw\n", column k as "%.{dec[k]}f"
struct RowFmt { int dec[8]; };
__device__ __forceinline__ int format_pose_row(const double* __restrict__ ts, const double* __restrict__ xyz, const double* __restrict__ quat,
                                               long long i, const RowFmt& F, char* buf, int* bad) {
    int len = 0;
    for (int k = 0; k < 8; ++k) {
        const double v = k == 0 ? ts[i] : (k < 4 ? xyz[3 * i + (k - 1)] : quat[4 * i + (k - 4)]);
        len += format_fixed(v, F.dec[k], buf + len, bad);
        buf[len++] = k == 7 ? '\n' : ' ';
    }
    return len;
}
constexpr int ROW_MAX = 8 * 32;
__global__ void __launch_bounds__(TX_T) fmt_lengths_kernel(const double* __restrict__ ts, const double* __restrict__ xyz, const double* __restrict__ quat,
                                                           long long n, RowFmt F, int* __restrict__ len, long long* __restrict__ counts, int* __restrict__ bad_out) {
    __shared__ long long tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    const long long i = (long long)blockIdx.x * TX_T + threadIdx.x;
    int l = 0;
    if (i < n) {
        char buf[ROW_MAX]; int bad = 0;
        l = format_pose_row(ts, xyz, quat, i, F, buf, &bad);
        len[i] = l;
        if (bad) *bad_out = 1;
    }
    long long s = l;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(GSF_FULL_MASK, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)&tot, (unsigned long long)s);
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(TX_T) fmt_write_kernel(const double* __restrict__ ts, const double* __restrict__ xyz, const double* __restrict__ quat,
                                                         long long n, RowFmt F, const int* __restrict__ len, const long long* __restrict__ counts,
                                                         long long header_bytes, char* __restrict__ out, long long capacity) {
    __shared__ long long wsum[TX_T / 32];
    const long long i = (long long)blockIdx.x * TX_T + threadIdx.x;
    const int l = i < n ? len[i] : 0;
    const long long off = header_bytes + counts[blockIdx.x] + block_exclusive_count((long long)l, 0, wsum);
    if (i < n && off + l <= capacity) {
        char buf[ROW_MAX]; int bad = 0;
        format_pose_row(ts, xyz, quat, i, F, buf, &bad);
        for (int k = 0; k < l; ++k) out[off + k] = buf[k];
    }
}
__global__ void fmt_finish_kernel(const long long* total, long long header_bytes, const int* bad, long long* out_info) {
    out_info[0] = header_bytes + *total; out_info[1] = *bad;
}

long long write_rows_work_bytes(long long n) { return 4 * ((n + 1) & ~1ll) + 8 * (div_up(n > 0 ? n : 1, TX_T) + 2) + 64; }
cudaError_t launch_write_pose_rows(const double* ts, const double* xyz, const double* quat, long long n, const int* decimals, const char* header,
                                   int header_bytes, char* out, long long capacity, void* work, long long* out_info, cudaStream_t stream) {
    RowFmt F;
    for (int k = 0; k < 8; ++k) { if (decimals[k] < 0 || decimals[k] > 9) return cudaErrorInvalidValue; F.dec[k] = decimals[k]; }
    const long long nb = div_up(n > 0 ? n : 1, TX_T);
    char* w = reinterpret_cast<char*>(work);
    int* len = reinterpret_cast<int*>(w); w += 4 * ((n + 1) & ~1ll);
    long long* counts = reinterpret_cast<long long*>(w); w += 8 * (nb + 1);
    long long* total = reinterpret_cast<long long*>(w); w += 8;
    int* bad = reinterpret_cast<int*>(w);
    cudaError_t e = cudaMemsetAsync(bad, 0, 8, stream);
    if (e != cudaSuccess) return e;
    if (header_bytes > 0) {
        if (header_bytes > capacity) return cudaErrorInvalidValue;
        e = cudaMemcpyAsync(out, header, (size_t)header_bytes, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
    }
    fmt_lengths_kernel<<<(unsigned)nb, TX_T, 0, stream>>>(ts, xyz, quat, n, F, len, counts, bad);
    scan_counts_kernel<<<1, 1024, 0, stream>>>(counts, nb, total);
    fmt_write_kernel<<<(unsigned)nb, TX_T, 0, stream>>>(ts, xyz, quat, n, F, len, counts, header_bytes, out, capacity);
    fmt_finish_kernel<<<1, 1, 0, stream>>>(total, header_bytes, bad, out_info);
    return cudaGetLastError();
}

}  // namespace gsf
