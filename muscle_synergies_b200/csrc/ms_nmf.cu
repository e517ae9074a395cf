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

// ---------------------------------------------------------------------------------------------------
// streaming regime: X and W too long for shared memory
// ---------------------------------------------------------------------------------------------------
// Per iteration and problem, X [n][m] and W [n][k] stream through HBM exactly once:
//   ms_nmf_stream_w_kernel   each CTA loops over 256-row tiles: tile of X and W -> shared memory
//                            (coalesced), W rows updated with the problem's H and H H^T, tile written
//                            back, and the tile's contribution to W^T X and W^T W accumulated in
//                            registers; one atomicAdd per CTA and output at the end
//   ms_nmf_stream_h_kernel   one small CTA per problem: H update, H H^T for the next iteration, the
//                            objective from ||X||^2 - 2 <H, W^T X> + <W^T W, H H^T> (no extra pass over X)
// Algorithmic bytes per iteration and problem: 4 n m (X) + 8 n k (W read + write).
#define NMFS_ROWS 256

struct MsNmfStreamState {  // per problem, device memory
    float HHt[NMF_MAX_K * NMF_MAX_K];
    float acc[NMF_MAX_K * NMF_MAX_K + NMF_MAX_K * NMF_MAX_M];  // W^T W then W^T X
    float err0, prev, err;
    int n_iter, done;
};

template <int K>
__device__ __forceinline__ void ms_nmf_stream_w_body(const float* __restrict__ X, long long n, int m, float* __restrict__ Wp,
                                                     const float* __restrict__ Hp, MsNmfStreamState* __restrict__ stt,
                                                     float* sm) {
    const int xs = m | 1;
    constexpr int ws = K | 1;
    float* sX = sm;                    // [NMFS_ROWS][xs]
    float* sW = sX + NMFS_ROWS * xs;   // [NMFS_ROWS][ws]
    float* sH = sW + NMFS_ROWS * ws;   // [K][m]
    float* sHHt = sH + K * m;          // [K][K]
    const int tid = threadIdx.x;
    for (int i = tid; i < K * m; i += NMFS_ROWS) sH[i] = Hp[i];
    for (int i = tid; i < K * K; i += NMFS_ROWS) sHHt[i] = stt->HHt[i];
    const int pairs = K * K + K * m;
    int split = NMFS_ROWS / pairs;
    if (split < 1) split = 1;
    const int rows_per = (NMFS_ROWS + split - 1) / split;
    float accum[(NMF_MAX_K * NMF_MAX_K + NMF_MAX_K * NMF_MAX_M + NMFS_ROWS - 1) / NMFS_ROWS];
#pragma unroll
    for (int q = 0; q < (int)(sizeof(accum) / sizeof(float)); q++) accum[q] = 0.f;
    const long long n_tiles = (n + NMFS_ROWS - 1) / NMFS_ROWS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long r0 = tile * NMFS_ROWS;
        const int rows = (int)min((long long)NMFS_ROWS, n - r0);
        __syncthreads();
        for (int i = tid; i < rows * m; i += NMFS_ROWS) sX[(i / m) * xs + (i % m)] = X[r0 * m + i];
        for (int i = tid; i < rows * K; i += NMFS_ROWS) sW[(i / K) * ws + (i % K)] = Wp[r0 * K + i];
        __syncthreads();
        if (tid < rows) {
            float w[K], num[K];
#pragma unroll
            for (int c = 0; c < K; c++) {
                w[c] = sW[tid * ws + c];
                num[c] = 0.f;
            }
            for (int j = 0; j < m; j++) {
                const float x = sX[tid * xs + j];
#pragma unroll
                for (int c = 0; c < K; c++) num[c] = fmaf(x, sH[c * m + j], num[c]);
            }
#pragma unroll
            for (int c = 0; c < K; c++) {
                float den = 0.f;
#pragma unroll
                for (int b = 0; b < K; b++) den = fmaf(w[b], sHHt[b * K + c], den);
                if (den == 0.f) den = NMF_EPS;
                sW[tid * ws + c] = w[c] * (num[c] / den);
            }
        }
        __syncthreads();
        for (int i = tid; i < rows * K; i += NMFS_ROWS) Wp[r0 * K + i] = sW[(i / K) * ws + (i % K)];
        int q = 0;
        for (int e = tid; e < pairs * split; e += NMFS_ROWS, q++) {
            const int pair = e % pairs, sl = e / pairs;
            const int i0 = sl * rows_per, i1 = min(rows, i0 + rows_per);
            float acc = 0.f;
            if (pair < K * K) {
                const int a = pair / K, b = pair % K;
                for (int i = i0; i < i1; i++) acc = fmaf(sW[i * ws + a], sW[i * ws + b], acc);
            } else {
                const int qq = pair - K * K, a = qq / m, j = qq % m;
                for (int i = i0; i < i1; i++) acc = fmaf(sW[i * ws + a], sX[i * xs + j], acc);
            }
            accum[q] += acc;
        }
    }
    int q = 0;
    for (int e = tid; e < pairs * split; e += NMFS_ROWS, q++) atomicAdd(&stt->acc[e % pairs], accum[q]);
}

__global__ void __launch_bounds__(NMFS_ROWS)
    ms_nmf_stream_w_kernel(const float* __restrict__ X, long long n, int m, const MsNmfProblem* __restrict__ problems,
                           float* __restrict__ Wg, const float* __restrict__ Hg, MsNmfStreamState* __restrict__ states) {
    extern __shared__ float sm[];
    const MsNmfProblem pb = problems[blockIdx.y];
    MsNmfStreamState* stt = states + blockIdx.y;
    if (stt->done) return;
    float* Wp = Wg + pb.w_off;
    const float* Hp = Hg + pb.h_off;
    const float* Xp = X + pb.x_off;
    switch (pb.k) {
#define MS_NMF_CASE(KK) \
    case KK:            \
        ms_nmf_stream_w_body<KK>(Xp, n, m, Wp, Hp, stt, sm); \
        break;
        MS_NMF_CASE(1) MS_NMF_CASE(2) MS_NMF_CASE(3) MS_NMF_CASE(4) MS_NMF_CASE(5) MS_NMF_CASE(6) MS_NMF_CASE(7)
        MS_NMF_CASE(8) MS_NMF_CASE(9) MS_NMF_CASE(10) MS_NMF_CASE(11) MS_NMF_CASE(12) MS_NMF_CASE(13)
        MS_NMF_CASE(14) MS_NMF_CASE(15) MS_NMF_CASE(16)
#undef MS_NMF_CASE
        default:
            break;
    }
}

// mode 0: prepare (H H^T of the initial H, clear accumulators); mode 1: H update after a W pass
__global__ void __launch_bounds__(128)
    ms_nmf_stream_h_kernel(int m, const MsNmfProblem* __restrict__ problems, float* __restrict__ Hg,
                           MsNmfStreamState* __restrict__ states, const float* __restrict__ xx, int mode, int iteration,
                           float tol, int check_every) {
    __shared__ float sH[NMF_MAX_K * NMF_MAX_M];
    const MsNmfProblem pb = problems[blockIdx.x];
    MsNmfStreamState* stt = states + blockIdx.x;
    const int k = pb.k, tid = threadIdx.x;
    float* Hp = Hg + pb.h_off;
    if (mode == 1 && stt->done) return;
    if (mode == 0) {
        if (tid == 0) {
            stt->done = 0;
            stt->n_iter = 0;
            stt->err0 = stt->prev = stt->err = 0.f;
        }
        for (int e = tid; e < k * m; e += 128) sH[e] = Hp[e];
    } else {
        const float* WtW = stt->acc;
        const float* WtX = stt->acc + k * k;
        for (int e = tid; e < k * m; e += 128) {
            const int a = e / m, j = e % m;
            float den = 0.f;
            for (int b = 0; b < k; b++) den = fmaf(WtW[a * k + b], Hp[b * m + j], den);
            if (den == 0.f) den = NMF_EPS;
            sH[e] = Hp[e] * (WtX[e] / den);
        }
    }
    __syncthreads();
    if (mode == 1)
        for (int e = tid; e < k * m; e += 128) Hp[e] = sH[e];
    for (int e = tid; e < k * k; e += 128) {
        const int a = e / k, b = e % k;
        float acc = 0.f;
        for (int j = 0; j < m; j++) acc = fmaf(sH[a * m + j], sH[b * m + j], acc);
        stt->HHt[e] = acc;
    }
    __syncthreads();
    if (mode == 1) {
        // ||X - W H||^2 = ||X||^2 - 2 <H, W^T X> + <W^T W, H H^T>   (W^T X, W^T W of the updated W)
        double part = 0.0;
        for (int e = tid; e < k * m; e += 128) part -= 2.0 * (double)sH[e] * (double)stt->acc[k * k + e];
        for (int e = tid; e < k * k; e += 128) part += (double)stt->acc[e] * (double)stt->HHt[e];
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        __shared__ double s_part[4];
        if ((tid & 31) == 0) s_part[tid >> 5] = part;
        __syncthreads();
        if (tid == 0) {
            const double sq = (double)xx[blockIdx.x] + s_part[0] + s_part[1] + s_part[2] + s_part[3];
            const float err = (float)sqrt(sq > 0.0 ? sq : 0.0);
            stt->err = err;
            stt->n_iter = iteration;
            if (iteration == 0) {
                stt->err0 = stt->prev = err;
            } else if (tol > 0.f && iteration % check_every == 0) {
                if ((stt->prev - err) / stt->err0 < tol) stt->done = 1;
                stt->prev = err;
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < k * k + k * m; e += 128) stt->acc[e] = 0.f;
}

// sum of squares of every column of X (per problem's X) and of the residual X - W H: VAF
__global__ void __launch_bounds__(256)
    ms_nmf_stream_vaf_kernel(const float* __restrict__ X, long long n, int m, const MsNmfProblem* __restrict__ problems,
                             const float* __restrict__ Wg, const float* __restrict__ Hg, double* __restrict__ sums) {
    // sums: [P][2][m]  (residual, total)
    __shared__ float sH[NMF_MAX_K * NMF_MAX_M];
    const MsNmfProblem pb = problems[blockIdx.y];
    const int k = pb.k;
    const float* Xp = X + pb.x_off;
    const float* Wp = Wg + pb.w_off;
    for (int e = threadIdx.x; e < k * m; e += 256) sH[e] = Hg[pb.h_off + e];
    __syncthreads();
    // thread -> column j = tid % m, rows strided: coalesced over the row-major X
    const int j = threadIdx.x % m, lane_row = threadIdx.x / m, rows_per_pass = 256 / m;
    if (lane_row >= rows_per_pass) return;
    double rs = 0.0, cs = 0.0;
    for (long long i = (long long)blockIdx.x * rows_per_pass + lane_row; i < n; i += (long long)gridDim.x * rows_per_pass) {
        float r = Xp[i * m + j];
        cs += (double)r * r;
        for (int c = 0; c < k; c++) r = fmaf(-Wp[i * k + c], sH[c * m + j], r);
        rs += (double)r * r;
    }
    atomicAdd(&sums[((long long)blockIdx.y * 2 + 0) * m + j], rs);
    atomicAdd(&sums[((long long)blockIdx.y * 2 + 1) * m + j], cs);
}

__global__ void ms_nmf_stream_finish_kernel(int m, const MsNmfStreamState* __restrict__ states,
                                            const double* __restrict__ sums, int32_t* __restrict__ n_iter,
                                            float* __restrict__ err, float* __restrict__ vaf, float* __restrict__ xx_out) {
    const int p = blockIdx.x;
    if (threadIdx.x == 0) {
        double rs = 0.0, cs = 0.0;
        for (int j = 0; j < m; j++) {
            rs += sums[((long long)p * 2 + 0) * m + j];
            cs += sums[((long long)p * 2 + 1) * m + j];
        }
        if (vaf) {
            vaf[(long long)p * (m + 1)] = (float)(1.0 - rs / cs);
            for (int j = 0; j < m; j++)
                vaf[(long long)p * (m + 1) + 1 + j] =
                    (float)(1.0 - sums[((long long)p * 2 + 0) * m + j] / sums[((long long)p * 2 + 1) * m + j]);
            n_iter[p] = states[p].n_iter;
            err[p] = (float)sqrt(rs);
        }
        if (xx_out) xx_out[p] = (float)cs;
    }
}

__global__ void ms_nmf_stream_err0_kernel(int m, MsNmfStreamState* __restrict__ states, const double* __restrict__ sums) {
    const int p = blockIdx.x;
    if (threadIdx.x == 0) {
        double rs = 0.0;
        for (int j = 0; j < m; j++) rs += sums[((long long)p * 2 + 0) * m + j];
        states[p].err0 = states[p].prev = states[p].err = (float)sqrt(rs);
    }
}

extern "C" int64_t ms_nmf_stream_workspace_bytes(int32_t m, int32_t n_problems) {
    return (int64_t)n_problems * ((int64_t)sizeof(MsNmfProblem) + (int64_t)sizeof(MsNmfStreamState) + 2 * m * 8 + 16) + 256;
}

// Same contract as ms_nmf_mu_batched for X of any length (X and W stream from HBM every
// iteration).  d_work: ms_nmf_stream_workspace_bytes(m, n_problems).  Synchronises `stream`
// every 64 iterations when tol > 0 to stop once every problem has converged.
extern "C" int ms_nmf_mu_stream(const float* d_X, int64_t n, int32_t m, const int32_t* h_ranks, const int32_t* h_x_index,
                                int32_t n_problems, float* d_W, float* d_H, int32_t max_iter, float tol,
                                int32_t check_every, void* d_work, int32_t* d_n_iter, float* d_err, float* d_vaf,
                                void* stream) {
    if (!d_X || !h_ranks || !d_W || !d_H || !d_work || !d_n_iter || !d_err || !d_vaf) return MS_E_INVALID;
    if (n < 1 || m < 1 || m > NMF_MAX_M || n_problems < 0 || max_iter < 0 || check_every < 1) return MS_E_INVALID;
    if (n_problems == 0) return MS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    MsNmfProblem* h = (MsNmfProblem*)malloc(sizeof(MsNmfProblem) * n_problems);
    if (!h) return MS_E_INVALID;
    long long wo = 0, ho = 0;
    int kmax = 0;
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
        wo += (long long)n * k;
        ho += (long long)k * m;
        if (k > kmax) kmax = k;
    }
    char* wsp = (char*)d_work;
    MsNmfProblem* d_problems = (MsNmfProblem*)wsp;
    wsp += ((sizeof(MsNmfProblem) * n_problems + 255) / 256) * 256;
    MsNmfStreamState* d_states = (MsNmfStreamState*)wsp;
    wsp += ((sizeof(MsNmfStreamState) * n_problems + 255) / 256) * 256;
    double* d_sums = (double*)wsp;
    wsp += (size_t)n_problems * 2 * m * 8;
    float* d_xx = (float*)wsp;
    cudaError_t e = cudaMemcpyAsync(d_problems, h, sizeof(MsNmfProblem) * n_problems, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    free(h);
    MS_CUDA_CHECK(e);

    const long long n_tiles = (n + NMFS_ROWS - 1) / NMFS_ROWS;
    int ctas = (148 * 4 + n_problems - 1) / n_problems;
    if (ctas > n_tiles) ctas = (int)n_tiles;
    if (ctas < 1) ctas = 1;
    const size_t smem = sizeof(float) * ((size_t)NMFS_ROWS * (m | 1) + (size_t)NMFS_ROWS * (kmax | 1) + (size_t)kmax * m +
                                         (size_t)kmax * kmax);
    MS_CUDA_CHECK(cudaFuncSetAttribute(ms_nmf_stream_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long vaf_want = 148 * 8 / (n_problems < 8 ? n_problems : 8) + 1, vaf_cap = (long long)((n + 15) / 16);
    int vaf_ctas = (int)(vaf_want < vaf_cap ? vaf_want : vaf_cap);

    // ||X||^2 per problem (for the objective), then the objective of the initial factors
    MS_CUDA_CHECK(cudaMemsetAsync(d_sums, 0, (size_t)n_problems * 2 * m * 8, st));
    ms_nmf_stream_vaf_kernel<<<dim3(vaf_ctas, n_problems), 256, 0, st>>>(d_X, n, m, d_problems, d_W, d_H, d_sums);
    MS_COUNT_LAUNCH();
    ms_nmf_stream_finish_kernel<<<n_problems, 32, 0, st>>>(m, d_states, d_sums, nullptr, nullptr, nullptr, d_xx);
    MS_COUNT_LAUNCH();
    ms_nmf_stream_h_kernel<<<n_problems, 128, 0, st>>>(m, d_problems, d_H, d_states, d_xx, 0, 0, tol, check_every);
    MS_COUNT_LAUNCH();
    // error_at_init: the residual sums just computed are those of (W0, H0)
    ms_nmf_stream_err0_kernel<<<n_problems, 32, 0, st>>>(m, d_states, d_sums);
    MS_COUNT_LAUNCH();

    for (int it = 1; it <= max_iter; it++) {
        ms_nmf_stream_w_kernel<<<dim3(ctas, n_problems), NMFS_ROWS, smem, st>>>(d_X, n, m, d_problems, d_W, d_H, d_states);
        MS_COUNT_LAUNCH();
        ms_nmf_stream_h_kernel<<<n_problems, 128, 0, st>>>(m, d_problems, d_H, d_states, d_xx, 1, it, tol, check_every);
        MS_COUNT_LAUNCH();
        if (tol > 0.f && it % 64 == 0) {
            // stop early when every problem has converged
            int all_done = 1;
            MsNmfStreamState* hs = (MsNmfStreamState*)malloc(sizeof(MsNmfStreamState) * n_problems);
            if (hs) {
                cudaMemcpyAsync(hs, d_states, sizeof(MsNmfStreamState) * n_problems, cudaMemcpyDeviceToHost, st);
                cudaStreamSynchronize(st);
                for (int p = 0; p < n_problems; p++)
                    if (!hs[p].done) all_done = 0;
                free(hs);
                if (all_done) break;
            }
        }
    }
    MS_CUDA_CHECK(cudaMemsetAsync(d_sums, 0, (size_t)n_problems * 2 * m * 8, st));
    ms_nmf_stream_vaf_kernel<<<dim3(vaf_ctas, n_problems), 256, 0, st>>>(d_X, n, m, d_problems, d_W, d_H, d_sums);
    MS_COUNT_LAUNCH();
    ms_nmf_stream_finish_kernel<<<n_problems, 32, 0, st>>>(m, d_states, d_sums, d_n_iter, d_err, d_vaf, nullptr);
    MS_COUNT_LAUNCH();
    MS_CUDA_CHECK(cudaGetLastError());
    return MS_OK;
}
