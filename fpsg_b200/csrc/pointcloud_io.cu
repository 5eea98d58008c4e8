// F4: the host input pipeline in front of the Chamfer path, as the reference's dataset runs it
// per item on the host (reference src/datasets/modelnet.py):
//   * ply_reader (:15-29): ASCII PLY -> vertex list                      -> fpsg_ply_read_ascii_host (host C++)
//   * FewShotSubModelNet.__getitem__ (:66-69): centre on the mean point,
//     divide by the largest point norm                                    -> fpsg_normalize_clouds (one CUDA kernel,
//                                                                            a whole batch of clouds per launch)
// The normalisation reproduces numpy's float32 arithmetic bit for bit: np.mean(axis=0) of a C-contiguous (N,3)
// float32 array adds the rows in order into three float32 accumulators and divides by N in float32; the norm is
// sqrt(rn(rn(x^2 + y^2) + z^2)); max; then (p - mean) / dist, every operation correctly rounded. The only
// sequential part is the mean (three 1-thread FP32 add chains per cloud, fed from shared memory); everything else
// is data-parallel.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace fpsg {

constexpr int kNormThreads = 256;
constexpr int kNormChunk = 2048;  // rows staged in shared memory per step of the sequential sums (24 KB)

__global__ void __launch_bounds__(kNormThreads) normalize_clouds_kernel(const float* __restrict__ in, int N, float* __restrict__ out) {
    __shared__ float sChunk[kNormChunk * 3];
    __shared__ float sMean[3];
    __shared__ float sRed[kNormThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float* __restrict__ src = in + (size_t)blockIdx.x * N * 3;
    float* __restrict__ dst = out + (size_t)blockIdx.x * N * 3;

    // mean: rows added in order, float32 accumulators (numpy's add.reduce over axis 0 of a C-contiguous array)
    float acc = 0.f;
    for (int r0 = 0; r0 < N; r0 += kNormChunk) {
        const int n = min(kNormChunk, N - r0);
        for (int e = tid; e < n * 3; e += kNormThreads) sChunk[e] = __ldg(src + (size_t)r0 * 3 + e);
        __syncthreads();
        if (tid < 3) {
#pragma unroll 8
            for (int r = 0; r < n; ++r) acc = __fadd_rn(acc, sChunk[3 * r + tid]);
        }
        __syncthreads();
    }
    if (tid < 3) sMean[tid] = __fdiv_rn(acc, (float)N);
    __syncthreads();
    const float mx = sMean[0], my = sMean[1], mz = sMean[2];

    // dist = max_i sqrt((x^2 + y^2) + z^2) of the centred points
    float best = 0.f;
    bool nan = false;
    for (int i = tid; i < N; i += kNormThreads) {
        const float cx = __fsub_rn(__ldg(src + 3 * (size_t)i), mx), cy = __fsub_rn(__ldg(src + 3 * (size_t)i + 1), my),
                    cz = __fsub_rn(__ldg(src + 3 * (size_t)i + 2), mz);
        const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
        nan |= d != d;
        best = fmaxf(best, d);
    }
    if (nan) best = __int_as_float(0x7fc00000);  // np.max propagates NaN
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float other = __shfl_xor_sync(0xffffffffu, best, o);
        best = (best != best || other != other) ? __int_as_float(0x7fc00000) : fmaxf(best, other);
    }
    if (lane == 0) sRed[w] = best;
    __syncthreads();
    float dist = sRed[0];
#pragma unroll
    for (int q = 1; q < kNormThreads / 32; ++q) dist = (dist != dist || sRed[q] != sRed[q]) ? __int_as_float(0x7fc00000) : fmaxf(dist, sRed[q]);

    for (int e = tid; e < N * 3; e += kNormThreads) {
        const int c = e % 3;
        dst[e] = __fdiv_rn(__fsub_rn(__ldg(src + e), c == 0 ? mx : (c == 1 ? my : mz)), dist);
    }
}

int normalize_clouds_launch(const float* in, int B, int N, float* out, cudaStream_t stream) {
    if (B < 0 || N < 1) return fail(FPSG_EINVAL, "normalize_clouds: need B >= 0, N >= 1 (got B=%d N=%d)", B, N);
    if (B == 0) return FPSG_OK;
    if (!in || !out) return fail(FPSG_EINVAL, "normalize_clouds: null pointer");
    if ((long long)N * 3 > 0x7fffffffLL) return fail(FPSG_EINVAL, "normalize_clouds: cloud too large");
    normalize_clouds_kernel<<<(unsigned)B, kNormThreads, 0, stream>>>(in, N, out);
    FPSG_CUDA_OK(cudaGetLastError());
    return FPSG_OK;
}

// ASCII PLY reader with the reference's semantics (src/datasets/modelnet.py:15-29): header lines are read until
// "end_header"; every header line with more than two space-separated tokens whose second token is "vertex" lowers
// the vertex count to min(count, max_points); then that many lines follow, each holding the coordinates of one
// vertex. Numbers go through strtod and are then rounded to float, exactly like Python's float() followed by
// np.asarray(..., dtype=np.float32). Lines must hold exactly three numbers (the reference keeps whatever it
// finds, but everything downstream of it needs (N,3)).
int ply_read_ascii(const char* path, int max_points, float* xyz, int* n_points) {
    if (!path || !xyz || !n_points || max_points < 1) return fail(FPSG_EINVAL, "ply_read_ascii: bad argument");
    FILE* f = fopen(path, "r");
    if (!f) return fail(FPSG_EINVAL, "ply_read_ascii: cannot open %s: %s", path, strerror(errno));
    char* line = nullptr;
    size_t cap = 0;
    int n_verts = max_points;
    bool header_done = false;
    while (getline(&line, &cap, f) >= 0) {
        // strip(): leading and trailing white space
        char* s = line;
        while (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\n' || *s == '\f' || *s == '\v') ++s;
        size_t len = strlen(s);
        while (len > 0 && (s[len - 1] == ' ' || s[len - 1] == '\t' || s[len - 1] == '\r' || s[len - 1] == '\n' || s[len - 1] == '\f' || s[len - 1] == '\v')) s[--len] = 0;
        if (strcmp(s, "end_header") == 0) {
            header_done = true;
            break;
        }
        // split(' '): tokens 1 and 2 (empty tokens between consecutive spaces count, as in Python)
        const char* t1 = strchr(s, ' ');
        if (!t1) continue;
        ++t1;
        const char* t2 = strchr(t1, ' ');
        if (!t2) continue;
        if ((size_t)(t2 - t1) == 6 && strncmp(t1, "vertex", 6) == 0) {
            ++t2;
            char* end = nullptr;
            const long v = strtol(t2, &end, 10);
            if (end == t2) {
                free(line);
                fclose(f);
                return fail(FPSG_EINVAL, "ply_read_ascii: %s: bad vertex count '%s'", path, t2);
            }
            if (v < n_verts) n_verts = (int)(v < 0 ? 0 : v);
        }
    }
    if (!header_done) {
        free(line);
        fclose(f);
        return fail(FPSG_EINVAL, "ply_read_ascii: %s: no end_header line", path);
    }
    for (int i = 0; i < n_verts; ++i) {
        if (getline(&line, &cap, f) < 0) {
            free(line);
            fclose(f);
            return fail(FPSG_EINVAL, "ply_read_ascii: %s: %d vertices announced, %d found", path, n_verts, i);
        }
        char* s = line;
        for (int c = 0; c < 3; ++c) {
            char* end = nullptr;
            const double v = strtod(s, &end);
            if (end == s) {
                free(line);
                fclose(f);
                return fail(FPSG_EINVAL, "ply_read_ascii: %s: vertex %d has fewer than 3 numbers", path, i);
            }
            xyz[3 * (size_t)i + c] = (float)v;
            s = end;
        }
        while (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\n') ++s;
        if (*s != 0) {
            free(line);
            fclose(f);
            return fail(FPSG_EINVAL, "ply_read_ascii: %s: vertex %d has more than 3 numbers (only xyz clouds are supported)", path, i);
        }
    }
    free(line);
    fclose(f);
    *n_points = n_verts;
    return FPSG_OK;
}

}  // namespace fpsg
