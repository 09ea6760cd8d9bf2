// Branch glue of one CFTM (ref M2Trans_network.py:135-161):
//   prep: n_k = InstanceNorm(X)[16k..16k+15] ; t_k = (n_k + y_{k-1})/2 (t_1 = n_1) ; Z = DWT^L(t_k)
//   post: y_k = IWT^L(O) + t_k                    with L = 0,1,2,2 for the four branches
// The InstanceNorm output is never materialised: (x - mu) * rstd is applied on load.
// Haar butterflies follow ref :203-209 (DWT) and :223-232 (IWT); band order LL,HL,LH,HH, channel index
// band*C + c at every level, so two levels give band2*64 + band1*16 + c.
// These are HBM/L2-bound elementwise kernels: 4 channels per thread, 16-byte loads of the fp32 stream.
#include "common.cuh"

namespace m2t {

struct F4 { float v[4]; };

__device__ __forceinline__ void haar_fwd(const F4& a, const F4& b, const F4& c, const F4& d, F4& ll, F4& hl,
                                         F4& lh, F4& hh) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        ll.v[e] = 0.5f * (a.v[e] + b.v[e] + c.v[e] + d.v[e]);
        hl.v[e] = 0.5f * (-a.v[e] - b.v[e] + c.v[e] + d.v[e]);
        lh.v[e] = 0.5f * (-a.v[e] + b.v[e] - c.v[e] + d.v[e]);
        hh.v[e] = 0.5f * (a.v[e] - b.v[e] - c.v[e] + d.v[e]);
    }
}
__device__ __forceinline__ void haar_inv(const F4& ll, const F4& hl, const F4& lh, const F4& hh, F4& a, F4& b,
                                         F4& c, F4& d) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        a.v[e] = 0.5f * (ll.v[e] - hl.v[e] - lh.v[e] + hh.v[e]);
        b.v[e] = 0.5f * (ll.v[e] - hl.v[e] + lh.v[e] - hh.v[e]);
        c.v[e] = 0.5f * (ll.v[e] + hl.v[e] - lh.v[e] - hh.v[e]);
        d.v[e] = 0.5f * (ll.v[e] + hl.v[e] + lh.v[e] + hh.v[e]);
    }
}

__device__ __forceinline__ F4 load_h4(const __half* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __half2 h0 = *reinterpret_cast<const __half2*>(&u.x);
    const __half2 h1 = *reinterpret_cast<const __half2*>(&u.y);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    F4 r; r.v[0] = f0.x; r.v[1] = f0.y; r.v[2] = f1.x; r.v[3] = f1.y;
    return r;
}
__device__ __forceinline__ void store_h4(__half* p, const F4& f) {
    const __half2 h0 = __floats2half2_rn(f.v[0], f.v[1]);
    const __half2 h1 = __floats2half2_rn(f.v[2], f.v[3]);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&h0);
    u.y = *reinterpret_cast<const uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(p) = u;
}

// t_k at full-resolution pixel `pix` (linear index into [B,Hp,Wp]) for 4 channels starting at 4q
__device__ __forceinline__ F4 load_t(const float* __restrict__ X, const __half* __restrict__ Y, long pix,
                                     int branch, int q, const float mu[4], const float rs[4]) {
    const float4 xv = *reinterpret_cast<const float4*>(X + pix * NF + NB * branch + 4 * q);
    F4 t;
    t.v[0] = (xv.x - mu[0]) * rs[0]; t.v[1] = (xv.y - mu[1]) * rs[1];
    t.v[2] = (xv.z - mu[2]) * rs[2]; t.v[3] = (xv.w - mu[3]) * rs[3];
    if (branch > 0) {
        const F4 yp = load_h4(Y + pix * NF + NB * (branch - 1) + 4 * q);
#pragma unroll
        for (int e = 0; e < 4; ++e) t.v[e] = (t.v[e] + yp.v[e]) * 0.5f;   // ref :141,:147,:155
    }
    return t;
}

// position index inside a 2x2 group: a=(0,0) b=(row+1) c=(col+1) d=(1,1)   (ref :204-207)
//   -> (dy,dx) = (p & 1, p >> 1)
template <int L>
__global__ void __launch_bounds__(128)
branch_prep_kernel(int branch, const float* __restrict__ X, const float2* __restrict__ munorm,
                   const __half* __restrict__ Y, __half* __restrict__ Z, int B, int Hp, int Wp) {
    constexpr int S = 1 << L;             // full-res pixels per level pixel, per axis
    constexpr int C = NB << (2 * L);      // channels at this level
    const int hl = Hp >> L, wl = Wp >> L;
    pdl_trigger();
    pdl_wait();
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)B * hl * wl * 4;
    if (tid >= total) return;
    const int q = (int)(tid & 3);
    const long lp = tid >> 2;
    const int b = (int)(lp / ((long)hl * wl));
    const int r = (int)(lp - (long)b * hl * wl);
    const int py = r / wl, px = r - py * wl;

    float mu[4], rs[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 m = __ldg(&munorm[b * NF + NB * branch + 4 * q + e]);
        mu[e] = m.x; rs[e] = m.y;
    }
    const long base = ((long)b * Hp + (long)py * S) * Wp + (long)px * S;
    __half* zo = Z + lp * C + 4 * q;
    if constexpr (L == 0) {
        store_h4(zo, load_t(X, Y, base, branch, q, mu, rs));
    } else if constexpr (L == 1) {
        F4 t[4], o[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) t[p] = load_t(X, Y, base + (long)(p & 1) * Wp + (p >> 1), branch, q, mu, rs);
        haar_fwd(t[0], t[1], t[2], t[3], o[0], o[1], o[2], o[3]);
#pragma unroll
        for (int bd = 0; bd < 4; ++bd) store_h4(zo + bd * NB, o[bd]);
    } else {
        F4 l1[4][4];   // [level-1 position][band1]
#pragma unroll
        for (int P = 0; P < 4; ++P) {
            F4 t[4];
            const long pb = base + (long)(2 * (P & 1)) * Wp + 2 * (P >> 1);
#pragma unroll
            for (int p = 0; p < 4; ++p) t[p] = load_t(X, Y, pb + (long)(p & 1) * Wp + (p >> 1), branch, q, mu, rs);
            haar_fwd(t[0], t[1], t[2], t[3], l1[P][0], l1[P][1], l1[P][2], l1[P][3]);
        }
#pragma unroll
        for (int b1 = 0; b1 < 4; ++b1) {
            F4 o[4];
            haar_fwd(l1[0][b1], l1[1][b1], l1[2][b1], l1[3][b1], o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) store_h4(zo + b2 * 4 * NB + b1 * NB, o[b2]);
        }
    }
}

template <int L>
__global__ void __launch_bounds__(128)
branch_post_kernel(int branch, const __half* __restrict__ O, const float* __restrict__ X,
                   const float2* __restrict__ munorm, __half* __restrict__ Y, int B, int Hp, int Wp) {
    constexpr int S = 1 << L;
    constexpr int C = NB << (2 * L);
    const int hl = Hp >> L, wl = Wp >> L;
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)B * hl * wl * 4;
    if (tid >= total) return;
    const int q = (int)(tid & 3);
    const long lp = tid >> 2;
    const int b = (int)(lp / ((long)hl * wl));
    const int r = (int)(lp - (long)b * hl * wl);
    const int py = r / wl, px = r - py * wl;

    float mu[4], rs[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 m = __ldg(&munorm[b * NF + NB * branch + 4 * q + e]);
        mu[e] = m.x; rs[e] = m.y;
    }
    const long base = ((long)b * Hp + (long)py * S) * Wp + (long)px * S;
    const __half* oi = O + lp * C + 4 * q;
    auto emit = [&](long pix, const F4& val) {
        const F4 t = load_t(X, Y, pix, branch, q, mu, rs);
        F4 yv;
#pragma unroll
        for (int e = 0; e < 4; ++e) yv.v[e] = val.v[e] + t.v[e];           // ref :139,:145,:153,:161
        store_h4(Y + pix * NF + NB * branch + 4 * q, yv);
    };
    if constexpr (L == 0) {
        emit(base, load_h4(oi));
    } else if constexpr (L == 1) {
        F4 o[4], t[4];
#pragma unroll
        for (int bd = 0; bd < 4; ++bd) o[bd] = load_h4(oi + bd * NB);
        haar_inv(o[0], o[1], o[2], o[3], t[0], t[1], t[2], t[3]);
#pragma unroll
        for (int p = 0; p < 4; ++p) emit(base + (long)(p & 1) * Wp + (p >> 1), t[p]);
    } else {
        F4 l1[4][4];   // [level-1 position][band1]
#pragma unroll
        for (int b1 = 0; b1 < 4; ++b1) {
            F4 o[4];
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) o[b2] = load_h4(oi + b2 * 4 * NB + b1 * NB);
            haar_inv(o[0], o[1], o[2], o[3], l1[0][b1], l1[1][b1], l1[2][b1], l1[3][b1]);
        }
#pragma unroll
        for (int P = 0; P < 4; ++P) {
            F4 t[4];
            haar_inv(l1[P][0], l1[P][1], l1[P][2], l1[P][3], t[0], t[1], t[2], t[3]);
            const long pb = base + (long)(2 * (P & 1)) * Wp + 2 * (P >> 1);
#pragma unroll
            for (int p = 0; p < 4; ++p) emit(pb + (long)(p & 1) * Wp + (p >> 1), t[p]);
        }
    }
}

// One pass over the residual stream at the start of a CFTM (ref :135-137, :141/:147/:155): every branch input
// that depends only on X is produced here, already in the layout its consumer wants:
//   T1 = n_1                      fp16 [B,Hp,Wp,16]                      (branch 1 input, t_1 = n_1)
//   H2 = n_2 / 2                  fp16 space-to-depth level 1 [B,Hp/2,Wp/2,64]
//   H3 = n_3 / 2, H4 = n_4 / 2    fp16 space-to-depth level 2 [B,Hp/4,Wp/4,256]
// The attention epilogue of branch k then completes t_{k+1} = H_{k+1} + y_k / 2 in place, reading and writing
// the same 32-byte segments, so it never touches the fp32 stream.
// One thread per (pixel, 16-channel branch group): 64-byte loads, 32-byte stores.  The InstanceNorm statistics
// (mean, rstd; biased variance, ref :127) are finalised here from the fp64 sums the producer accumulated, so no
// separate finalise pass is needed on this path.  Persistent CTAs, each with a contiguous range of 64-pixel chunks:
// the fp64 finalise (divide + square root on 64 threads, ~1.5 K cycles) runs once per CTA and image instead of once
// per 64 pixels, where it cost as much as the data movement.
// Measured and not kept (round 2): the next chunk's loads issued before the current chunk is converted, coordinates carried
// along instead of two divisions per chunk, the residual computed only where it is stored, one wave of 4 CTAs per SM:
// 24.8 vs 26.6 us per launch at cfg2, but 626 vs 599 us at cfg4 and 154 vs 145 us at cfg3, where this kernel already runs
// at 87 % of the copy bandwidth.  The same kernel with 8 CTAs per SM: 26.7 / 154 / 607 us (no better anywhere); what the
// one-wave form saves at cfg2 is the per-CTA fp64 finalise of the statistics (half as many CTAs).
__global__ void __launch_bounds__(256)
branch_prep_all_kernel(const float* __restrict__ X, const double* __restrict__ stats, __half* __restrict__ T1,
                       __half* __restrict__ T1lo, __half* __restrict__ H2, __half* __restrict__ H3,
                       __half* __restrict__ H4, int B, int Hp, int Wp, int resident) {
    __shared__ float smu[NF], srs[NF];
    // the CTAs of the last (on small inputs: the only) wave let the next kernel's launch proceed now, so that its CTAs start
    // the moment ours leave (the register file is full, they cannot take our slots early)
    pdl_trigger_last_wave(resident);
    pdl_wait();
    const int t = threadIdx.x;
    const int npix = Hp * Wp;
    const int nchunks = (int)((long)B * npix / 64);
    const int k_lo = (int)((long)blockIdx.x * nchunks / gridDim.x), k_hi = (int)((long)(blockIdx.x + 1) * nchunks / gridDim.x);
    const int branch = t & 3;
    const float sc = branch == 0 ? 1.f : 0.5f;
    int cur_b = -1;
    for (int k = k_lo; k < k_hi; ++k) {
        // The producer (head or the previous ff conv) wrote X front to back just before this kernel: walk it back to
        // front so the most recently written part is read while it is still in L2.
        const long pix = (long)(nchunks - 1 - k) * 64 + (t >> 2);
        const int b = (int)(pix / npix);
        const int r = (int)(pix - (long)b * npix);
        const int y = r / Wp, x = r - y * Wp;
        const float4* xp = reinterpret_cast<const float4*>(X + pix * NF + NB * branch);
        uint4 a[4];                        // 64 contiguous bytes per thread as two 256-bit loads
        ldg256(xp, a[0], a[1]);
        ldg256(xp + 2, a[2], a[3]);
        if (b != cur_b) {                  // uniform over the CTA: a chunk never straddles two images
            __syncthreads();
            if (t < NF) {
                const double inv = 1.0 / (double)npix;
                const double m = stats[((long)b * NF + t) * 2] * inv;
                double var = stats[((long)b * NF + t) * 2 + 1] * inv - m * m;
                if (var < 0.0) var = 0.0;
                smu[t] = (float)m;
                srs[t] = (float)(1.0 / sqrt(var + (double)IN_EPS));
            }
            __syncthreads();
            cur_b = b;
        }
        uint4 o[2], ol[2];                 // fp16 value and (branch 1 only) its rounding residual
        __half2* oh = reinterpret_cast<__half2*>(o);
        __half2* olh = reinterpret_cast<__half2*>(ol);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int c = NB * branch + 4 * v;
            const float n0 = (__uint_as_float(a[v].x) - smu[c]) * srs[c] * sc, n1 = (__uint_as_float(a[v].y) - smu[c + 1]) * srs[c + 1] * sc;
            const float n2 = (__uint_as_float(a[v].z) - smu[c + 2]) * srs[c + 2] * sc, n3 = (__uint_as_float(a[v].w) - smu[c + 3]) * srs[c + 3] * sc;
            oh[2 * v] = __floats2half2_rn(n0, n1);
            oh[2 * v + 1] = __floats2half2_rn(n2, n3);
            const float2 h01 = __half22float2(oh[2 * v]), h23 = __half22float2(oh[2 * v + 1]);
            olh[2 * v] = __floats2half2_rn(n0 - h01.x, n1 - h01.y);
            olh[2 * v + 1] = __floats2half2_rn(n2 - h23.x, n3 - h23.y);
        }
        __half* dst;
        if (branch == 0) {
            dst = T1 + pix * NB;
            if (T1lo != nullptr) stg256(T1lo + pix * NB, ol[0], ol[1]);
        } else if (branch == 1) {
            const long lp = ((long)b * (Hp >> 1) + (y >> 1)) * (Wp >> 1) + (x >> 1);
            dst = H2 + lp * 64 + ((y & 1) * 2 + (x & 1)) * NB;
        } else {
            const long lp = ((long)b * (Hp >> 2) + (y >> 2)) * (Wp >> 2) + (x >> 2);
            dst = (branch == 2 ? H3 : H4) + lp * 256 + ((y & 3) * 4 + (x & 3)) * NB;
        }
        stg256(dst, o[0], o[1]);
    }
}

int launch_branch_prep_all(const float* X, const double* stats, __half* T1, __half* T1lo, __half* H2, __half* H3,
                           __half* H4, const Geom& g, cudaStream_t s) {
    const long nchunks = (long)g.B * g.Hp * g.Wp / 64;
    // 256 threads x 64 registers: 4 CTAs per SM are resident.  Every CTA finalises the statistics of its image(s) in fp64
    // (~1.5 K cycles): with few chunks per CTA one wave of CTAs beats two half as long ones (cfg2: 24.4 vs 26.3 us), with many
    // chunks per CTA the finer split balances better (cfg4: 599 vs 626 us).
    const long sms = device_sm_count();
#ifndef PREP_SMALL_MULT
#define PREP_SMALL_MULT 4
#endif
    const long slots = (nchunks < 32L * 4L * sms ? (long)PREP_SMALL_MULT : 8L) * sms;
    const unsigned grid = (unsigned)(nchunks < slots ? nchunks : slots);
    M2T_CUDA(launch_pdl(branch_prep_all_kernel, dim3(grid), dim3(256), 0, s, X, stats, T1, T1lo, H2, H3, H4, g.B, g.Hp, g.Wp,
                        resident_ctas(branch_prep_all_kernel, 256, 0)));
    return M2T_OK;
}

int launch_branch_prep(int level, int branch, const float* X, const float2* munorm, const __half* Y, __half* Z,
                       const Geom& g, cudaStream_t s) {
    const long total = (long)g.B * (g.Hp >> level) * (g.Wp >> level) * 4;
    const unsigned grid = (unsigned)((total + 127) / 128);
    if (level == 0) M2T_CUDA(launch_pdl(branch_prep_kernel<0>, dim3(grid), dim3(128), 0, s, branch, X, munorm, Y, Z, g.B, g.Hp, g.Wp));
    else if (level == 1) M2T_CUDA(launch_pdl(branch_prep_kernel<1>, dim3(grid), dim3(128), 0, s, branch, X, munorm, Y, Z, g.B, g.Hp, g.Wp));
    else M2T_CUDA(launch_pdl(branch_prep_kernel<2>, dim3(grid), dim3(128), 0, s, branch, X, munorm, Y, Z, g.B, g.Hp, g.Wp));
    return M2T_OK;
}

int launch_branch_post(int level, int branch, const __half* O, const float* X, const float2* munorm, __half* Y,
                       const Geom& g, cudaStream_t s) {
    const long total = (long)g.B * (g.Hp >> level) * (g.Wp >> level) * 4;
    const unsigned grid = (unsigned)((total + 127) / 128);
    if (level == 0) branch_post_kernel<0><<<grid, 128, 0, s>>>(branch, O, X, munorm, Y, g.B, g.Hp, g.Wp);
    else if (level == 1) branch_post_kernel<1><<<grid, 128, 0, s>>>(branch, O, X, munorm, Y, g.B, g.Hp, g.Wp);
    else branch_post_kernel<2><<<grid, 128, 0, s>>>(branch, O, X, munorm, Y, g.B, g.Hp, g.Wp);
    M2T_LAUNCH_CHECK("branch_post_kernel");
    return M2T_OK;
}

}  // namespace m2t
