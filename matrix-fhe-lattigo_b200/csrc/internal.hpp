// internal.hpp -- structures shared by the C-ABI layer and the kernel launchers.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b200ring.h"
#include "modarith.cuh"
#include "solinas.cuh"

// {w, ws}: a constant and its Shoup companion floor(w*2^64/q); one 128-bit load.
typedef ulonglong2 tw_t;

// host copy of one prime's constants (ring.SubRing + NTTTable, subring.go:35-55, ntt.go:38-44)
struct SubRingHost {
    u64 q = 0, qinv = 0, brc[2] = {0, 0}, mask = 0;
    u64 g = 0;        // SubRing.PrimitiveRoot (smallest primitive root >= 3, or caller supplied)
    u64 ninv = 0;     // NTTTable.NInv, Montgomery form (STANDARD)
    u64 root = 0;     // psi = g^((q-1)/2N) (STANDARD) or omega (MATRIX), plain domain
    std::vector<u64> roots_fwd, roots_bwd;   // Go-format tables (STANDARD): Montgomery, bit-reversed
};

// indices into the per-limb misc constant block (MATRIX rings)
enum { MISC_OMEGA3 = 0, MISC_K1 = 1, MISC_K2 = 2, MISC_ZETA = 3, MISC_NINV = 4, MISC_NINV_W1 = 5, MISC_COUNT = 8 };
// conjugate-invariant rings: the fourth roots of unity of the folding layer (RootsForward[1], RootsBackward[1])
enum { MISC_CI_F = 6, MISC_CI_FINV = 7 };
// MATRIX rings: N^-1 (the inverse cyclotomic layer in its two-product form, ntt_fold.cuh) shares the slot of MISC_NINV
enum { MISC_K4 = 4 };
// The inverse scaling constants times 2^64 mod q, for the fused ring product on the fold path: its slot-wise product is a
// Montgomery product (x*y*2^-64, 9 wide multiplies instead of Barrett's 18) and the inverse transform that follows gives the
// factor back for free.  MATRIX: slots 5, 6; STANDARD: slots 6, 7 (conjugate-invariant rings, which own 6 and 7, do not take
// the fused path).  Selected in the inverse column kernels by the internal flag B200_FLAG_MONT_SCALE.
enum { MISC_K1M = 5, MISC_K4M = 6, MISC_NINVM = 6, MISC_NINV_W1M = 7 };
#define B200_FLAG_MONT_SCALE 0x100

struct b200_ctx {
    int ring_type = 0, N = 0, L = 0, device = 0;
    int core_type = 0;     // butterfly network the kernels run: STANDARD (also for conjugate-invariant rings) or MATRIX
    int logN = 0;          // STANDARD: log2 N
    int a = 0, b = 0;      // MATRIX: N = 2^a 3^b
    cudaStream_t stream = nullptr;
    int max_smem_optin = 0, sm_count = 0;
    std::vector<SubRingHost> sub;
    LimbConst *d_consts = nullptr;   // [L]
    tw_t *d_fwd = nullptr;           // [L][N] forward twiddles (STANDARD: psi^brev(k); MATRIX: zetas)
    tw_t *d_inv = nullptr;           // [L][N] their inverses
    tw_t *d_misc = nullptr;          // [L][MISC_COUNT]
    tw_t *d_fwd_r = nullptr, *d_inv_r = nullptr;   // rings the two-pass kernels cover: copies of d_fwd / d_inv whose entries of heap index >= N/256
                                     // (all the row pass reads) sit in row-pass order (row_pass_order, b200ring.cu)
    // fold arithmetic (solinas.cuh): set when every prime is 2^k + c with 50 <= k <= 60, |c| <= 2^(k-33) and every twiddle
    // splits into 32-bit halves; the tuned transforms then run on these tables instead of the Shoup ones
    bool fold_ok = false;
    SolConst *d_sol = nullptr;       // [L]
    stw_t *d_fwd_s = nullptr, *d_inv_s = nullptr, *d_misc_s = nullptr;   // same indexing as d_fwd / d_inv / d_misc
    int *d_go2fast = nullptr;        // [N] MATRIX: reference slot k -> tree slot i
    u64 *d_scratch = nullptr;        // grown on demand (generic large-N path, polymul temporaries)
    size_t scratch_words = 0;
    u64 *d_ws = nullptr;             // second workspace for composite operators that call the transforms (rescale)
    size_t ws_words = 0;
    std::vector<u64> rescale;        // RescaleConstants[level-1][i] (ring.go:361-379), row-major [L-1][L], Montgomery form
    u64 *d_rescale = nullptr;        // the same table on the device
    u64 *d_index = nullptr;          // staging for caller-supplied automorphism tables (AutomorphismNTTWithIndex)
    size_t index_words = 0;
    std::atomic<uint64_t> launches{0};
    // Ring/SubRing are "thread safe" in the reference (ring/ring.go:192-194): goroutines may call the same ring
    // concurrently.  Every compute entry point holds this lock while it enqueues, so the launches of one call
    // (which share the scratch buffers) are never interleaved with another call's on the context's stream.
    std::recursive_mutex mu;
    // optional per-kernel device timing (b200_profile_begin/end): CUDA events around each launch
    bool profiling = false;
    struct ProfRec { const char *name; cudaEvent_t e0, e1; long long units; };
    std::vector<ProfRec> prof;
    // optimised-path tables (ntt_big.cu); null when the ring shape is not covered
    void *big = nullptr;
    // pipelined host-buffer path (b200_transform_host): copy/compute streams, events and staging buffers
    struct HostPipe {
        cudaStream_t h2d = nullptr, comp = nullptr, d2h = nullptr;
        static constexpr int NB = 3;
        u64 *buf[NB] = {nullptr, nullptr, nullptr};
        size_t buf_words = 0;
        cudaEvent_t up[NB], done[NB], freed[NB];
        cudaEvent_t entry = nullptr;   // orders the pipeline's compute stream after the work already queued on the context's stream
        bool ready = false;
    } pipe;
};

struct b200_poly {
    b200_ctx *ctx = nullptr;
    u64 *d = nullptr;
    int batch = 0;
    bool owned = false;
};

struct CtxLock {
    std::recursive_mutex *m;
    explicit CtxLock(const b200_ctx *c) : m(c ? &const_cast<b200_ctx *>(c)->mu : nullptr) { if (m) m->lock(); }
    ~CtxLock() { if (m) m->unlock(); }
    CtxLock(const CtxLock &) = delete;
    CtxLock &operator=(const CtxLock &) = delete;
};

void b200_set_error(const std::string &msg);
int b200_cuda_fail(cudaError_t e, const char *what);
#define CUDA_TRY(expr)                                                    \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess) return b200_cuda_fail(_e, #expr);          \
    } while (0)

int b200_ensure_scratch(b200_ctx *ctx, size_t words);
int b200_ensure_scratch2(b200_ctx *ctx, size_t words);
// the stream launchers enqueue on: the context's stream, or a per-thread override used by the pipelined
// host-buffer entry point (b200_transform_host) while it drives its own compute stream
cudaStream_t b200_cur_stream(b200_ctx *ctx);
void b200_set_stream_override(cudaStream_t s, bool active);

// RAII bracket around one kernel launch: counts it and, when profiling, times it with CUDA events
struct LaunchScope {
    b200_ctx *c;
    int idx = -1;
    LaunchScope(b200_ctx *ctx, const char *name, long long units) : c(ctx) {
        c->launches++;
        if (c->profiling) {
            b200_ctx::ProfRec r{name, nullptr, nullptr, units};
            cudaEventCreate(&r.e0);
            cudaEventCreate(&r.e1);
            cudaEventRecord(r.e0, b200_cur_stream(c));
            c->prof.push_back(r);
            idx = (int)c->prof.size() - 1;
        }
    }
    ~LaunchScope() {
        if (idx >= 0) cudaEventRecord(c->prof[idx].e1, b200_cur_stream(c));
    }
};

// kernel launchers (each returns 0 or a negative code; all enqueue on ctx->stream)
// A launch covers `batch` polynomials x limbs [limb0, limb0+nl) of the ring; the buffers are laid out
// [batch][Lbuf][N] and buffer limb j (j < nl) uses the ring's prime limb0 + j.
struct Span {
    int batch, nl, Lbuf, limb0;
};
int launch_vec_op(b200_ctx *ctx, int op, const u64 *p1, const u64 *p2, u64 *p3, Span sp, const u64 *d_scalars);
int launch_div_last(b200_ctx *ctx, const u64 *in, u64 *out, int batch, int level, bool round);
int launch_mul_by_vector(b200_ctx *ctx, const u64 *p1, const u64 *vector, u64 *p3, Span sp, bool then_add_lazy);
int launch_rescale_spread(b200_ctx *ctx, u64 *ws, int batch, int level, bool round);
int launch_tensor(b200_ctx *ctx, const u64 *a0, const u64 *a1, const u64 *b0, const u64 *b1, u64 *c0, u64 *c1, u64 *c2,
                  Span sp, bool mform_first);
int launch_ntt_generic(b200_ctx *ctx, const u64 *in, u64 *out, Span sp, bool inverse, int flags);
// optimised two-pass path; returns 1 if the shape is not covered (caller falls back to generic)
int big_tables_create(b200_ctx *ctx);
void big_tables_destroy(b200_ctx *ctx);
// Extra destinations of an inverse transform on other GPUs of the box (peer device memory mapped through CUDA IPC):
// the last pass stores every output word to `out` and to the same offset of each p[i] (fused all-gather over NVLink).
struct PeerOut {
    int n = 0;
    u64 *p[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
int launch_ntt_big(b200_ctx *ctx, const u64 *in, u64 *out, Span sp, bool inverse, int flags, const PeerOut *peers = nullptr);
// the transform of the ring, whatever its type: tuned two-pass kernels where the shape is covered, generic kernels
// otherwise, plus the folding layer of conjugate-invariant rings (ring/ntt.go:754-786,1104-1158)
int launch_ntt(b200_ctx *ctx, const u64 *in, u64 *out, Span sp, bool inverse, int flags);
int launch_ci_fold(b200_ctx *ctx, const u64 *in, u64 *out, Span sp, bool inverse);
// ring/automorphism.go: NTT-domain gather (mode 0 assign, 1 add, 2 fold standard -> conjugate-invariant) and the
// coefficient-domain permutation with sign
int launch_automorphism_ntt(b200_ctx *ctx, const u64 *in, u64 *out, const u64 *d_index, u64 gen, u64 nthroot, int n_in, int l_in,
                            int batch, int nl, int mode);
int launch_automorphism_coeff(b200_ctx *ctx, const u64 *in, u64 *out, u64 gen, int batch, int nl);
int launch_rotate(b200_ctx *ctx, const u64 *in, u64 *out, int batch, int nl, int s, int neg, int mode);
int launch_ci_unfold(b200_ctx *ctx, const u64 *in, u64 *out, int N, int batch, int nl, int Lin, int Lout);
int launch_polymul_big(b200_ctx *ctx, const u64 *a, const u64 *b, u64 *c, Span sp);
