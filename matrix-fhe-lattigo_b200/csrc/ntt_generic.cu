// ntt_generic.cu -- shape-agnostic transform kernels: any power-of-two N and any N = 2^a 3^b.
//
// Two drivers share the same per-layer device functions:
//   * "resident" kernel: one CTA per (poly, limb) unit, the whole limb staged in shared memory,
//     all layers in ONE launch with one HBM read and one HBM write (N*8 <= 227 KB, i.e. the
//     reference's test sizes, the KAT sizes and BASELINE config 1, N = 3*2^12);
//   * "layered" kernels: one launch per layer over global memory, for shapes that neither fit in
//     shared memory nor are covered by the tuned two-pass kernels of ntt_big.cu (e.g. b >= 2 at
//     large N).  Correct for every valid ring; not the performance path.
//
// Algorithms: STANDARD = Cooley-Tukey forward / Gentleman-Sande inverse with merged psi powers,
// i.e. what ring/ntt.go:223-257,568-606 computes, with Harvey lazy butterflies on Shoup twiddles.
// MATRIX = cyclotomic layer + b radix-3 layers + (a-1) radix-2 layers (the
// structure of the reference prototype references/integer_dft.py:266-432), slots permuted to the
// reference order of ring/ntt_3n.go:46-53.
#include "internal.hpp"

namespace {

struct Shape {
    int N, logN, a, b;
    Span sp;
};
__device__ __forceinline__ size_t unit_base(const Shape &s, int unit, int &limb) {
    limb = unit % s.sp.nl;
    return ((size_t)(unit / s.sp.nl) * s.sp.Lbuf + limb) * (size_t)s.N;
}

// ---- STANDARD layers ----------------------------------------------------------------------------
// forward layer m (m blocks of 2t): in [0,4q) -> out [0,4q)
__device__ __forceinline__ void std_fwd_layer(u64 *d, int half, int logt, int m, const tw_t *tw, u64 q, int tid, int nth) {
    const u64 twoq = 2 * q;
    const int t = 1 << logt;
    for (int j = tid; j < half; j += nth) {
        const int i = j >> logt, off = j & (t - 1);
        const int x = (i << (logt + 1)) + off;
        const tw_t w = tw[m + i];
        const u64 U = csub(d[x], twoq);
        const u64 T = shoup_lazy(d[x + t], w.x, w.y, q);
        d[x] = U + T;
        d[x + t] = U - T + twoq;
    }
}
// inverse layer with h blocks of 2t: in [0,2q) -> out [0,2q)
__device__ __forceinline__ void std_inv_layer(u64 *d, int half, int logt, int h, const tw_t *tw, u64 q, int tid, int nth) {
    const u64 twoq = 2 * q;
    const int t = 1 << logt;
    for (int j = tid; j < half; j += nth) {
        const int i = j >> logt, off = j & (t - 1);
        const int x = (i << (logt + 1)) + off;
        const tw_t w = tw[h + i];
        const u64 U = d[x], V = d[x + t];
        d[x] = csub(U + V, twoq);
        d[x + t] = shoup_lazy(U - V + twoq, w.x, w.y, q);
    }
}
// last inverse layer (t = N/2) with N^-1 folded in: out canonical (or [0,2q) when lazy)
__device__ __forceinline__ void std_inv_last(u64 *d, int half, const tw_t ninv, const tw_t ninv_w1, u64 q, bool lazy, int tid, int nth) {
    const u64 twoq = 2 * q;
    for (int j = tid; j < half; j += nth) {
        const u64 U = d[j], V = d[j + half];
        u64 X = shoup_lazy(U + V, ninv.x, ninv.y, q);
        u64 Y = shoup_lazy(U - V + twoq, ninv_w1.x, ninv_w1.y, q);
        if (!lazy) { X = csub(X, q); Y = csub(Y, q); }
        d[j] = X;
        d[j + half] = Y;
    }
}

// ---- MATRIX layers ------------------------------------------------------------------------------
// cyclotomic layer: in [0,2q) -> out [0,4q)   (lo' = lo + z hi, hi' = lo + hi - z hi)
__device__ __forceinline__ void m3_fwd_cyc(u64 *d, int N, const tw_t zeta, u64 q, int tid, int nth) {
    const int half = N >> 1;
    const u64 twoq = 2 * q, fourq = 4 * q;
    for (int j = tid; j < half; j += nth) {
        const u64 lo = d[j], hi = d[j + half];
        const u64 t = shoup_lazy(hi, zeta.x, zeta.y, q);
        d[j] = lo + t;
        d[j + half] = csub(lo + hi + twoq - t, fourq);
    }
}
// radix-3 layer with `blocks` blocks of 3*step: in [0,4q) -> out [0,4q)
__device__ __forceinline__ void m3_fwd_r3(u64 *d, int N, int step, int blocks, const tw_t *tw, const tw_t om, u64 q, int tid, int nth) {
    const u64 twoq = 2 * q, fourq = 4 * q;
    const int third = N / 3;
    for (int j = tid; j < third; j += nth) {
        const int blk = j / step, off = j - blk * step;
        const int x = blk * 3 * step + off;
        const tw_t z1 = tw[blocks + 2 * blk], z2 = tw[blocks + 2 * blk + 1];
        const u64 b0 = csub(d[x], twoq);
        const u64 t1 = shoup_lazy(d[x + step], z1.x, z1.y, q);
        const u64 t2 = shoup_lazy(d[x + 2 * step], z2.x, z2.y, q);
        const u64 t3 = shoup_lazy(t1 + twoq - t2, om.x, om.y, q);
        d[x] = csub(b0 + t1 + t2, fourq);                       // < 6q -> < 4q
        d[x + step] = csub(b0 + twoq - t2 + t3, fourq);         // < 6q -> < 4q
        d[x + 2 * step] = csub(b0 + fourq - t1 - t3, fourq);    // < 6q -> < 4q
    }
}
// radix-2 layer with `blocks` blocks of 2*step (same butterfly as the STANDARD forward layer)
__device__ __forceinline__ void m3_fwd_r2(u64 *d, int N, int logstep, int blocks, const tw_t *tw, u64 q, int tid, int nth) {
    std_fwd_layer(d, N >> 1, logstep, blocks, tw, q, tid, nth);
}
// inverse radix-2 layer: in [0,2q) -> out [0,2q), twiddles = zeta^-1
__device__ __forceinline__ void m3_inv_r2(u64 *d, int N, int logstep, int blocks, const tw_t *twi, u64 q, int tid, int nth) {
    std_inv_layer(d, N >> 1, logstep, blocks, twi, q, tid, nth);
}
// inverse radix-3 layer (factor 3 deferred): in [0,2q) -> out [0,2q)
__device__ __forceinline__ void m3_inv_r3(u64 *d, int N, int step, int blocks, const tw_t *twi, const tw_t om, u64 q, int tid, int nth) {
    const u64 twoq = 2 * q, fourq = 4 * q;
    const int third = N / 3;
    for (int j = tid; j < third; j += nth) {
        const int blk = j / step, off = j - blk * step;
        const int x = blk * 3 * step + off;
        const tw_t z1 = twi[blocks + 2 * blk], z2 = twi[blocks + 2 * blk + 1];
        const u64 y0 = d[x], y1 = d[x + step], y2 = d[x + 2 * step];
        const u64 s = shoup_lazy(y1 + twoq - y2, om.x, om.y, q);           // Omega (y1 - y2)
        d[x] = csub(csub(y0 + y1 + y2, fourq), twoq);
        d[x + step] = shoup_lazy(y0 + fourq - y1 - s, z1.x, z1.y, q);      // z1^-1 (y0 - y1 - s)
        d[x + 2 * step] = shoup_lazy(y0 + twoq - y2 + s, z2.x, z2.y, q);   // z2^-1 (y0 - y2 + s)
    }
}
// inverse cyclotomic layer with the 2/N scaling: hi = K1 (lo' - hi'), lo = K2 lo' - zeta hi
__device__ __forceinline__ void m3_inv_cyc(u64 *d, int N, const tw_t *misc, u64 q, bool lazy, int tid, int nth) {
    const int half = N >> 1;
    const u64 twoq = 2 * q;
    const tw_t k1 = misc[MISC_K1], k2 = misc[MISC_K2], zeta = misc[MISC_ZETA];
    for (int j = tid; j < half; j += nth) {
        const u64 lo = d[j], hi = d[j + half];
        u64 h = shoup_lazy(lo + twoq - hi, k1.x, k1.y, q);
        u64 l = shoup_lazy(lo, k2.x, k2.y, q) + twoq - shoup_lazy(h, zeta.x, zeta.y, q);
        l = csub(l, twoq);
        if (!lazy) { l = csub(l, q); h = csub(h, q); }
        d[j] = l;
        d[j + half] = h;
    }
}

// All layers of one unit on data d (shared or global), `sync()` between layers.
template <bool INV, typename Sync>
__device__ __forceinline__ void run_layers(u64 *d, const Shape &s, int ring_type, const tw_t *fwd, const tw_t *inv,
                                           const tw_t *misc, u64 q, bool lazy, int tid, int nth, Sync sync) {
    const int N = s.N;
    if (ring_type == B200_RING_STANDARD) {
        if (!INV) {
            for (int m = 1, logt = s.logN - 1; m < N; m <<= 1, --logt) {
                std_fwd_layer(d, N >> 1, logt, m, fwd, q, tid, nth);
                sync();
            }
        } else {
            for (int h = N >> 1, logt = 0; h > 1; h >>= 1, ++logt) {
                std_inv_layer(d, N >> 1, logt, h, inv, q, tid, nth);
                sync();
            }
            std_inv_last(d, N >> 1, misc[MISC_NINV], misc[MISC_NINV_W1], q, lazy, tid, nth);
            sync();
        }
    } else {
        const int r2_first = 1 << (s.a - 1);   // smallest radix-3 step
        if (!INV) {
            m3_fwd_cyc(d, N, fwd[1], q, tid, nth);
            sync();
            for (int step = N / 6, blocks = 2; step >= r2_first; step /= 3, blocks *= 3) {
                m3_fwd_r3(d, N, step, blocks, fwd, misc[MISC_OMEGA3], q, tid, nth);
                sync();
            }
            for (int logstep = s.a - 2; logstep >= 0; --logstep) {
                m3_fwd_r2(d, N, logstep, N >> (logstep + 1), fwd, q, tid, nth);
                sync();
            }
        } else {
            for (int logstep = 0; logstep <= s.a - 2; ++logstep) {
                m3_inv_r2(d, N, logstep, N >> (logstep + 1), inv, q, tid, nth);
                sync();
            }
            for (int step = r2_first; step <= N / 6; step *= 3) {
                m3_inv_r3(d, N, step, N / (3 * step), inv, misc[MISC_OMEGA3], q, tid, nth);
                sync();
            }
            m3_inv_cyc(d, N, misc, q, lazy, tid, nth);
            sync();
        }
    }
}

// ---- resident kernel: one CTA per unit, limb in shared memory ----------------------------------------
template <bool INV>
__global__ void __launch_bounds__(1024)
k_ntt_resident(const u64 *in, u64 *out, const LimbConst *__restrict__ consts, const tw_t *__restrict__ fwd,
               const tw_t *__restrict__ inv, const tw_t *__restrict__ misc, const int *__restrict__ go2fast,
               Shape s, int ring_type, int flags) {
    extern __shared__ __align__(16) u64 sm[];
    int limb;
    const size_t base = unit_base(s, blockIdx.x, limb);
    limb += s.sp.limb0;
    const u64 q = consts[limb].q;
    const int tid = threadIdx.x, nth = blockDim.x, N = s.N;
    const bool lazy = flags & B200_LAZY;
    const bool permute = ring_type == B200_RING_MATRIX && !(flags & B200_TREE_ORDER);
    if (INV && permute) {
        for (int k = tid; k < N; k += nth) sm[go2fast[k]] = in[base + k];
    } else {
        for (int k = tid; k < N; k += nth) sm[k] = in[base + k];
    }
    __syncthreads();
    run_layers<INV>(sm, s, ring_type, fwd + (size_t)limb * N, inv + (size_t)limb * N, misc + limb * MISC_COUNT, q, lazy,
                    tid, nth, [] { __syncthreads(); });
    if (!INV) {
        const u64 twoq = 2 * q;
        for (int k = tid; k < N; k += nth) {
            u64 v = sm[permute ? go2fast[k] : k];
            if (!lazy) v = csub(csub(v, twoq), q);   // [0,4q) -> canonical; lazy leaves [0,4q) (<= 6q-2)
            out[base + k] = v;
        }
    } else {
        for (int k = tid; k < N; k += nth) out[base + k] = sm[k];
    }
}

// ---- layered kernels: one launch per layer, grid = (chunks, units) ---------------------------------
struct LayerArgs {
    int kind;      // 0 std fwd, 1 std inv, 2 std inv last, 3 cyc fwd, 4 r3 fwd, 5 r2 fwd, 6 r2 inv, 7 r3 inv, 8 cyc inv
    int p0, p1;    // (logt|step|logstep), (m|h|blocks)
};
__global__ void __launch_bounds__(256)
k_ntt_layer(u64 *data, const LimbConst *__restrict__ consts, const tw_t *__restrict__ fwd, const tw_t *__restrict__ inv,
            const tw_t *__restrict__ misc, Shape s, LayerArgs la, int flags, int unit0) {
    int limb;
    u64 *d = data + unit_base(s, unit0 + blockIdx.y, limb);
    limb += s.sp.limb0;
    const u64 q = consts[limb].q;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x, N = s.N;
    const tw_t *f = fwd + (size_t)limb * N, *iv = inv + (size_t)limb * N, *mc = misc + limb * MISC_COUNT;
    const bool lazy = flags & B200_LAZY;
    switch (la.kind) {
    case 0: std_fwd_layer(d, N >> 1, la.p0, la.p1, f, q, tid, nth); break;
    case 1: std_inv_layer(d, N >> 1, la.p0, la.p1, iv, q, tid, nth); break;
    case 2: std_inv_last(d, N >> 1, mc[MISC_NINV], mc[MISC_NINV_W1], q, lazy, tid, nth); break;
    case 3: m3_fwd_cyc(d, N, f[1], q, tid, nth); break;
    case 4: m3_fwd_r3(d, N, la.p0, la.p1, f, mc[MISC_OMEGA3], q, tid, nth); break;
    case 5: m3_fwd_r2(d, N, la.p0, la.p1, f, q, tid, nth); break;
    case 6: m3_inv_r2(d, N, la.p0, la.p1, iv, q, tid, nth); break;
    case 7: m3_inv_r3(d, N, la.p0, la.p1, iv, mc[MISC_OMEGA3], q, tid, nth); break;
    case 8: m3_inv_cyc(d, N, mc, q, lazy, tid, nth); break;
    }
}
// copy / slot permutation / final reduction between the caller's buffers and the work buffer
// mode 0: dst[k] = src[k]; 1: dst[go2fast[k]] = src[k] (inverse input); 2: dst[k] = reduce(src[go2fast[k] or k]) (forward output)
__global__ void __launch_bounds__(256)
k_ntt_move(const u64 *src, u64 *dst, const LimbConst *__restrict__ consts, const int *__restrict__ go2fast, Shape s,
           int mode, int permute, int lazy, int unit0) {
    int limb;
    const size_t base = unit_base(s, unit0 + blockIdx.y, limb);
    const u64 q = consts[s.sp.limb0 + limb].q;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < s.N; k += gridDim.x * blockDim.x) {
        if (mode == 0) dst[base + k] = src[base + k];
        else if (mode == 1) dst[base + go2fast[k]] = src[base + k];
        else {
            u64 v = src[base + (permute ? go2fast[k] : k)];
            if (!lazy) v = csub(csub(v, 2 * q), q);
            dst[base + k] = v;
        }
    }
}

}  // namespace

int launch_ntt_generic(b200_ctx *ctx, const u64 *in, u64 *out, Span sp, bool inverse, int flags) {
    Shape s{ctx->N, ctx->logN, ctx->a, ctx->b, sp};
    const int N = ctx->N, units = sp.batch * sp.nl;
    const size_t smem = (size_t)N * sizeof(u64);
    if (smem <= (size_t)ctx->max_smem_optin) {
        int threads = N / 2 >= 1024 ? 1024 : (N / 2 < 32 ? 32 : ((N / 2 + 31) / 32) * 32);
        if (threads > 1024) threads = 1024;
        auto kern = inverse ? k_ntt_resident<true> : k_ntt_resident<false>;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            LaunchScope ls(ctx, inverse ? "resident_inv" : "resident_fwd", units);
            kern<<<units, threads, smem, b200_cur_stream(ctx)>>>(in, out, ctx->d_consts, ctx->d_fwd, ctx->d_inv, ctx->d_misc,
                                                        ctx->d_go2fast, s, ctx->core_type, flags);
        }
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    // layered path: work in place on `out` (tree order), with a scratch copy when a slot permutation is needed
    const bool matrix = ctx->ring_type == B200_RING_MATRIX;
    const bool permute = matrix && !(flags & B200_TREE_ORDER);
    const int lazy = (flags & B200_LAZY) ? 1 : 0;
    const size_t words = (size_t)sp.batch * sp.Lbuf * N;
    u64 *work = out;
    int chunks = (N / 2 + 255) / 256;
    if (chunks > 512) chunks = 512;
    auto for_units = [&](auto &&fn) {
        for (int u0 = 0; u0 < units; u0 += 65535) fn(u0, units - u0 < 65535 ? units - u0 : 65535);
    };
    auto move = [&](const u64 *src, u64 *dst, int mode, int perm) {
        for_units([&](int u0, int nu) {
            LaunchScope ls(ctx, "layered_move", nu);
            k_ntt_move<<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(src, dst, ctx->d_consts, ctx->d_go2fast, s, mode, perm, lazy, u0);
        });
    };
    auto layer = [&](int kind, int p0, int p1) {
        LayerArgs la{kind, p0, p1};
        for_units([&](int u0, int nu) {
            LaunchScope ls(ctx, "layered_layer", nu);
            k_ntt_layer<<<dim3(chunks, nu), 256, 0, b200_cur_stream(ctx)>>>(work, ctx->d_consts, ctx->d_fwd, ctx->d_inv, ctx->d_misc, s, la, flags, u0);
        });
    };
    if (!inverse) {
        if (permute) {   // transform in scratch, gather into out at the end
            int rc = b200_ensure_scratch(ctx, words);
            if (rc) return rc;
            work = ctx->d_scratch;
            move(in, work, 0, 0);
        } else if (in != out) {
            move(in, out, 0, 0);
        }
        if (!matrix) {
            for (int m = 1, logt = ctx->logN - 1; m < N; m <<= 1, --logt) layer(0, logt, m);
        } else {
            layer(3, 0, 0);
            for (int step = N / 6, blocks = 2; step >= (1 << (ctx->a - 1)); step /= 3, blocks *= 3) layer(4, step, blocks);
            for (int logstep = ctx->a - 2; logstep >= 0; --logstep) layer(5, logstep, N >> (logstep + 1));
        }
        move(work, out, 2, permute ? 1 : 0);   // final reduction (+ slot permutation)
    } else {
        if (permute) {
            if (in == out) {   // scatter needs a distinct source
                int rc = b200_ensure_scratch(ctx, words);
                if (rc) return rc;
                move(in, ctx->d_scratch, 0, 0);
                move(ctx->d_scratch, out, 1, 1);
            } else {
                move(in, out, 1, 1);
            }
        } else if (in != out) {
            move(in, out, 0, 0);
        }
        if (!matrix) {
            for (int h = N >> 1, logt = 0; h > 1; h >>= 1, ++logt) layer(1, logt, h);
            layer(2, 0, 0);
        } else {
            for (int logstep = 0; logstep <= ctx->a - 2; ++logstep) layer(6, logstep, N >> (logstep + 1));
            for (int step = 1 << (ctx->a - 1); step <= N / 6; step *= 3) layer(7, step, N / (3 * step));
            layer(8, 0, 0);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
