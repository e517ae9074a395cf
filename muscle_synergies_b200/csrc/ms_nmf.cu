// Batched NMF by multiplicative updates (Frobenius loss) - EXTENSION stage.
//
// The reference only wraps scikit-learn here: find_synergies -> NMF(n_components=k, **kwargs)
// .fit_transform(X) (src/muscle_synergies/analysis.py:848-882) and vaf (:597-667).  This file
// implements, for solver="mu", beta_loss="frobenius", the update sklearn executes
// (sklearn/decomposition/_nmf.py, _fit_multiplicative_update):
//
//     W <- W * (X H^T) / (W (H H^T))      denominators == 0 replaced by EPS (float32 eps)
//     H <- H * (W^T X) / ((W^T W) H)      with the UPDATED W
//     every `check_every` iterations (tol > 0): e = ||X - W H||_F ; stop if (prev - e) / e0 < tol
//
// in fp32 for a whole batch of problems (rank sweep x random restarts) per launch.  It is
// checked against sklearn within a stated tolerance (tests/test_nmf_gpu.py); it is not a
// bit-parity claim.
//
//   ms_nmf_resident_kernel   one CTA per problem, X / W / H resident in shared memory for the
//                            whole run (the 200 x 16 envelopes of the reference flow are 12.8 KB):
//                            no HBM traffic between iterations, bound by SM issue + shared memory
//   ms_nmf_stream_*          long signals: X and W stream from HBM once per iteration, W^T X and
//                            W^T W reduced per CTA and accumulated with atomics (HBM-bound)
#include <stdio.h>

#include "ms_common.cuh"

#define NMF_EPS 1.1920929e-07f
#define NMF_THREADS 256
#define NMF_MAX_K 16
#define NMF_MAX_M 64

struct MsNmfProblem {
    int k;
    long long w_off;  // floats from d_W to this problem's W [n][k]
    long long h_off;  // floats from d_H to this problem's H [k][m]
    long long x_off;  // floats from d_X to this problem's X [n][m]
};

__device__ __forceinline__ float ms_block_sum(float v, float* s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_red[w];
    return t;
}

// ---------------------------------------------------------------------------------------------------
// resident regime
// ---------------------------------------------------------------------------------------------------
struct MsNmfArgs {
    const float* X;
    int n, m;
    float* Wp;
    float* Hp;
    int max_iter;
    float tol;
    int check_every;
    int32_t* n_iter_out;
    float* err_out;
    float* vaf_out;  // [m + 1] of this problem
};

template <int K>
__device__ __forceinline__ void ms_nmf_resident_body(const MsNmfArgs& A, float* sm) {
    const int n = A.n, m = A.m;
    const int xs = m | 1;  // odd row strides: consecutive rows hit different banks
    constexpr int ws = K | 1;
    float* sX = sm;                // [n][xs]
    float* sW = sX + n * xs;       // [n][ws]
    float* sH = sW + n * ws;       // [K][m]
    float* sHHt = sH + K * m;      // [K][K]
    float* sWtW = sHHt + K * K;    // [K][K]
    float* sWtX = sWtW + K * K;    // [K][m]   (contiguous with sWtW)
    float* s_red = sWtX + K * m;   // [32]
    const int tid = threadIdx.x;

    for (int i = tid; i < n * m; i += NMF_THREADS) sX[(i / m) * xs + (i % m)] = A.X[i];
    for (int i = tid; i < n * K; i += NMF_THREADS) sW[(i / K) * ws + (i % K)] = A.Wp[i];
    for (int i = tid; i < K * m; i += NMF_THREADS) sH[i] = A.Hp[i];
    __syncthreads();

    auto residual_sq = [&]() {
        float acc = 0.f;
        for (int i = tid; i < n; i += NMF_THREADS) {
            float w[K];
#pragma unroll
            for (int c = 0; c < K; c++) w[c] = sW[i * ws + c];
            for (int j = 0; j < m; j++) {
                float r = sX[i * xs + j];
#pragma unroll
                for (int c = 0; c < K; c++) r = fmaf(-w[c], sH[c * m + j], r);
                acc = fmaf(r, r, acc);
            }
        }
        return ms_block_sum(acc, s_red);
    };
    auto compute_hht = [&]() {
        for (int e = tid; e < K * K; e += NMF_THREADS) {
            const int a = e / K, b = e % K;
            float acc = 0.f;
            for (int j = 0; j < m; j++) acc = fmaf(sH[a * m + j], sH[b * m + j], acc);
            sHHt[e] = acc;
        }
    };

    float err0 = 0.f, prev = 0.f;
    if (A.tol > 0.f) {
        err0 = sqrtf(residual_sq());
        prev = err0;
    }
    compute_hht();
    __syncthreads();

    int it = 0;
    for (it = 1; it <= A.max_iter; it++) {
        // ---- W <- W * (X H^T) / (W (H H^T)); zero the accumulators of the H step meanwhile
        for (int e = tid; e < K * K + K * m; e += NMF_THREADS) sWtW[e] = 0.f;
        for (int i = tid; i < n; i += NMF_THREADS) {
            float w[K], num[K];
#pragma unroll
            for (int c = 0; c < K; c++) {
                w[c] = sW[i * ws + c];
                num[c] = 0.f;
            }
            for (int j = 0; j < m; j++) {
                const float x = sX[i * xs + j];
#pragma unroll
                for (int c = 0; c < K; c++) num[c] = fmaf(x, sH[c * m + j], num[c]);
            }
#pragma unroll
            for (int c = 0; c < K; c++) {
                float den = 0.f;
#pragma unroll
                for (int b = 0; b < K; b++) den = fmaf(w[b], sHHt[b * K + c], den);
                if (den == 0.f) den = NMF_EPS;
                sW[i * ws + c] = w[c] * (num[c] / den);
            }
        }
        __syncthreads();
        // ---- W^T W and W^T X: slices of rows per thread group, reduced with shared atomics
        {
            const int pairs = K * K + K * m;
            int split = NMF_THREADS / pairs;
            if (split < 1) split = 1;
            const int rows_per = (n + split - 1) / split;
            for (int e = tid; e < pairs * split; e += NMF_THREADS) {
                const int pair = e % pairs, sl = e / pairs;
                const int i0 = sl * rows_per, i1 = min(n, i0 + rows_per);
                float acc = 0.f;
                if (pair < K * K) {
                    const int a = pair / K, b = pair % K;
                    for (int i = i0; i < i1; i++) acc = fmaf(sW[i * ws + a], sW[i * ws + b], acc);
                } else {
                    const int q = pair - K * K, a = q / m, j = q % m;
                    for (int i = i0; i < i1; i++) acc = fmaf(sW[i * ws + a], sX[i * xs + j], acc);
                }
                atomicAdd(&sWtW[pair], acc);
            }
        }
        __syncthreads();
        // ---- H <- H * (W^T X) / ((W^T W) H)
        {
            float newh[(NMF_MAX_K * NMF_MAX_M + NMF_THREADS - 1) / NMF_THREADS];
            int q = 0;
            for (int e = tid; e < K * m; e += NMF_THREADS, q++) {
                const int a = e / m, j = e % m;
                float den = 0.f;
#pragma unroll
                for (int b = 0; b < K; b++) den = fmaf(sWtW[a * K + b], sH[b * m + j], den);
                if (den == 0.f) den = NMF_EPS;
                newh[q] = sH[e] * (sWtX[e] / den);
            }
            __syncthreads();
            q = 0;
            for (int e = tid; e < K * m; e += NMF_THREADS, q++) sH[e] = newh[q];
        }
        __syncthreads();
        compute_hht();
        __syncthreads();
        if (A.tol > 0.f && it % A.check_every == 0) {
            const float err = sqrtf(residual_sq());
            if ((prev - err) / err0 < A.tol) break;
            prev = err;
        }
    }
    if (it > A.max_iter) it = A.max_iter;

    // ---- results: factors, ||X - W H||_F, variance accounted for (analysis.py:642-667)
    for (int i = tid; i < n * K; i += NMF_THREADS) A.Wp[i] = sW[(i / K) * ws + (i % K)];
    for (int i = tid; i < K * m; i += NMF_THREADS) A.Hp[i] = sH[i];
    const float res = residual_sq();
    float xx = 0.f;
    for (int i = tid; i < n * m; i += NMF_THREADS) {
        const float v = sX[(i / m) * xs + (i % m)];
        xx = fmaf(v, v, xx);
    }
    xx = ms_block_sum(xx, s_red);
    if (tid == 0) {
        *A.n_iter_out = it;
        *A.err_out = sqrtf(res);
        A.vaf_out[0] = 1.f - res / xx;
    }
    // per-muscle VAF: one column per thread
    for (int j = tid; j < m; j += NMF_THREADS) {
        float rs = 0.f, cs = 0.f;
        for (int i = 0; i < n; i++) {
            float r = sX[i * xs + j];
            cs = fmaf(r, r, cs);
#pragma unroll
            for (int c = 0; c < K; c++) r = fmaf(-sW[i * ws + c], sH[c * m + j], r);
            rs = fmaf(r, r, rs);
        }
        A.vaf_out[1 + j] = 1.f - rs / cs;
    }
}

__global__ void __launch_bounds__(NMF_THREADS)
    ms_nmf_resident_kernel(const float* __restrict__ X, int n, int m, const MsNmfProblem* __restrict__ problems,
                           float* __restrict__ Wg, float* __restrict__ Hg, int max_iter, float tol, int check_every,
                           int32_t* __restrict__ n_iter_out, float* __restrict__ err_out, float* __restrict__ vaf_out) {
    extern __shared__ float sm[];
    const MsNmfProblem pb = problems[blockIdx.x];
    MsNmfArgs A;
    A.X = X + pb.x_off;
    A.n = n;
    A.m = m;
    A.Wp = Wg + pb.w_off;
    A.Hp = Hg + pb.h_off;
    A.max_iter = max_iter;
    A.tol = tol;
    A.check_every = check_every;
    A.n_iter_out = n_iter_out + blockIdx.x;
    A.err_out = err_out + blockIdx.x;
    A.vaf_out = vaf_out + (long long)blockIdx.x * (m + 1);
    switch (pb.k) {
#define MS_NMF_CASE(KK) \
    case KK:            \
        ms_nmf_resident_body<KK>(A, sm); \
        break;
        MS_NMF_CASE(1) MS_NMF_CASE(2) MS_NMF_CASE(3) MS_NMF_CASE(4) MS_NMF_CASE(5) MS_NMF_CASE(6) MS_NMF_CASE(7)
        MS_NMF_CASE(8) MS_NMF_CASE(9) MS_NMF_CASE(10) MS_NMF_CASE(11) MS_NMF_CASE(12) MS_NMF_CASE(13)
        MS_NMF_CASE(14) MS_NMF_CASE(15) MS_NMF_CASE(16)
#undef MS_NMF_CASE
        default:
            break;
    }
}

static size_t ms_nmf_resident_smem(int n, int m, int kmax) {
    const int xs = m | 1, ws = kmax | 1;
    return sizeof(float) * ((size_t)n * xs + (size_t)n * ws + 2 * (size_t)kmax * m + 2 * (size_t)kmax * kmax + 32);
}

// Largest n the resident kernel takes for (m, kmax); 0 if the shape is unsupported.
extern "C" int32_t ms_nmf_resident_max_rows(int32_t m, int32_t kmax) {
    if (m < 1 || m > NMF_MAX_M || kmax < 1 || kmax > NMF_MAX_K) return 0;
    const size_t budget = 227 * 1024 - 1024;
    const size_t fixed = sizeof(float) * (2 * (size_t)kmax * m + 2 * (size_t)kmax * kmax + 32);
    const size_t per_row = sizeof(float) * ((size_t)(m | 1) + (size_t)(kmax | 1));
    return (int32_t)((budget - fixed) / per_row);
}

// h_ranks[P]: rank of each problem; h_x_index[P] (may be NULL = all 0): which [n][m] matrix of d_X
// problem p factorises.  d_W / d_H hold the initial factors packed problem after problem (W_p is
// [n][k_p] row-major, H_p is [k_p][m]) and receive the results in place.
// d_work: P * 32 bytes.  d_vaf: [P][m + 1] (overall, then per column).
extern "C" int ms_nmf_mu_batched(const float* d_X, int32_t n, int32_t m, const int32_t* h_ranks,
                                 const int32_t* h_x_index, int32_t n_problems,
                                 float* d_W, float* d_H, int32_t max_iter, float tol, int32_t check_every, void* d_work,
                                 int32_t* d_n_iter, float* d_err, float* d_vaf, void* stream) {
    if (!d_X || !h_ranks || !d_W || !d_H || !d_work || !d_n_iter || !d_err || !d_vaf) return MS_E_INVALID;
    if (n < 1 || m < 1 || m > NMF_MAX_M || n_problems < 0 || max_iter < 0 || check_every < 1) return MS_E_INVALID;
    if (n_problems == 0) return MS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int kmax = 0;
    MsNmfProblem* h = (MsNmfProblem*)malloc(sizeof(MsNmfProblem) * n_problems);
    if (!h) return MS_E_INVALID;
    long long wo = 0, ho = 0;
    for (int p = 0; p < n_problems; p++) {
        const int k = h_ranks[p];
        if (k < 1 || k > NMF_MAX_K) {
            free(h);
            return MS_E_INVALID;
        }
        h[p].k = k;
        h[p].w_off = wo;
        h[p].h_off = ho;
        h[p].x_off = h_x_index ? (long long)h_x_index[p] * n * m : 0;
        if (h[p].x_off < 0) {
            free(h);
            return MS_E_INVALID;
        }
        wo += (long long)n * k;
        ho += (long long)k * m;
        if (k > kmax) kmax = k;
    }
    if (n > ms_nmf_resident_max_rows(m, kmax)) {
        free(h);
        return MS_E_INVALID;  // too long for the resident kernel: use ms_nmf_mu_stream
    }
    cudaError_t e = cudaMemcpyAsync(d_work, h, sizeof(MsNmfProblem) * n_problems, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // h is pageable: do not free it under the copy
    free(h);
    MS_CUDA_CHECK(e);
    const size_t smem = ms_nmf_resident_smem(n, m, kmax);
    MS_CUDA_CHECK(cudaFuncSetAttribute(ms_nmf_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ms_nmf_resident_kernel<<<n_problems, NMF_THREADS, smem, st>>>(d_X, n, m, (const MsNmfProblem*)d_work, d_W, d_H,
                                                                  max_iter, tol, check_every, d_n_iter, d_err, d_vaf);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}
