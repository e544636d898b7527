// Per-edge MLP of ML3Layer on the tensor cores (reference libs/spect_conv.py:191-194 and :206-207):
//     ea' = relu( W4 [ relu(W1 ea) || tanh(W2 ea) * tanh(W3 ea) ] )         W1,W2,W3: [2K, K], W4: [K, 4K], no bias
// The CUDA-core version (edge_mlp.cu) spends 640 (forward) / 1920 (backward) FP32 FMAs per edge with every weight read
// as a shared-memory broadcast: 1312 issued instructions per edge forward, issue-bound at 37 % of the FP32 peak, 27 % of
// the ZINC training step.  Here the four small contractions of an edge run as tcgen05.mma on tiles of 128 edges:
//   * thread = edge = TENSOR-MEMORY LANE.  A worker group of 128 threads owns a tile; every operand that depends on the edge
//     (ea, the hidden activations, d pre4, d pre1..3) is written by its thread into its own lane with tcgen05.st
//     (32x32b: one lane, consecutive columns) and read by the tensor core as the A operand from tensor memory; the
//     results come back with tcgen05.ld into the same thread.  Nothing edge-dependent crosses threads, so no shared
//     memory staging, no transposes and no bank conflicts are involved in the forward pass.
//   * the weights are the B operands: K-major SWIZZLE_128B planes built once per CTA in shared memory, rows = output
//     features with the TF32 hi parts in the first half and the residuals lo = w - hi in the second half.  Three
//     instructions per 8-wide k-step,  D += A_raw W_hi^T,  D += A_raw W_lo^T,  D += A_lo W_hi^T,  accumulate the
//     error-compensated 3xTF32 product in ONE set of columns (FP32-grade; the tensor core truncates the raw FP32 operand to
//     its hi part itself).  The tensor pipe is 15 % busy, instruction issue is what bounds the kernel: summing the
//     partial products in the accumulator instead of in registers removed a third of the tcgen05.ld traffic and 72 FADDs
//     per edge.
//   * edge-feature widths are padded to P = 8 or 16 (k-steps of 8), the three first-layer branches sit at plane rows
//     0 / 2P / 4P, so that every tensor-memory region is a whole number of 16-column blocks and the activations can be
//     computed block by block IN PLACE (the block of hidden activations overwrites exactly the pre-activation columns it
//     was computed from; column map in k_edge_mlp_tc).
//   * the MMAs of a group are issued by one elected thread of the group itself (named barrier -> tcgen05.mma ->
//     tcgen05.commit -> mbarrier), up to four groups per CTA run their tiles independently and hide each other's
//     round trips.
// Backward: recompute (same instructions as the forward, so the ReLU masks agree bit for bit), d tmp = d pre4 W4 and
// (when d ea is wanted) d ea = d pre123 W123 on the tensor core; the four weight gradients -- contractions over the EDGE
// axis, i.e. over tensor-memory lanes -- stay on the FP32 pipe as register-tiled outer products over a per-group staging
// tile (the phase-2 scheme of edge_mlp.cu), now overlapped with the other groups' tensor-core phases.  Per-group
// partials are reduced in a fixed order by k_edge_mlp_bwd_reduce: deterministic, no atomics.
#include "edge_mlp.cuh"
#include "tc_common.cuh"

namespace gnnml3 {

#ifndef EMT_MAXWG
#define EMT_MAXWG 4
#endif

template <int K>
struct EMT {
    using C = EMC<K>;
    static constexpr int P = (K + 7) / 8 * 8;           // padded edge-feature width: 8 or 16
    static constexpr int HP = 2 * P, TP = 4 * P, DPP = 6 * P;
    static constexpr int N1 = 6 * P;                    // first layer: [W1;W2;W3] rows (hi rows, then as many lo rows)
    static constexpr int N2 = 16;                       // second layer and d ea: W4 / W123^T rows
    static constexpr int N3 = TP;                       // d tmp = d pre4 W4: W4^T rows
    static constexpr int TB = TP / 32;                  // 128-byte column blocks of the second layer's contraction
    static constexpr int DB = (DPP + 31) / 32;          // ... of the d ea contraction over the 6P pre-activation gradients
    static constexpr int B1_BYTES = 2 * N1 * 128, B2_BYTES = TB * 32 * 128, B3_BYTES = 2 * N3 * 128, B4_BYTES = DB * 32 * 128;
    // tensor-memory columns of one worker group: X = pre123 (6P) -> hidden raw | lo (8P) -> d tmp (4P) -> d pre123 raw | lo (12P);
    // Y = ea raw | lo (2P) -> pre4 (16) -> d pre4 raw | lo (2P) -> d ea (16)
    __host__ __device__ static constexpr int xc(int mode) { return mode > 1 ? 12 * P : 8 * P; }
    static constexpr int YC = 2 * P;
    __host__ __device__ static constexpr int slot(int mode) { return xc(mode) + YC; }
    __host__ __device__ static constexpr int plane_bytes(int mode) { return B1_BYTES + B2_BYTES + (mode > 0 ? B3_BYTES : 0) + (mode > 1 ? B4_BYTES : 0); }
    static constexpr int stage_bytes = 128 * C::S * 4;
    static constexpr int SMEM_MAX = 227 * 1024 - 1024;  // minus the 1024-byte alignment slack
    __host__ __device__ static constexpr int nwg(int mode) {
        int n = 512 / slot(mode);
        if (mode > 0) {
            const int m = (SMEM_MAX - plane_bytes(mode) - 64) / stage_bytes;
            n = m < n ? m : n;
        }
        return n < 1 ? 1 : (n > EMT_MAXWG ? EMT_MAXWG : n);
    }
    __host__ __device__ static constexpr int smem_bytes(int mode) { return 1024 + plane_bytes(mode) + 64 + (mode > 0 ? nwg(mode) * stage_bytes : 0); }
    // phase 2 (weight gradients) on 128 threads
    static constexpr int NT = C::NT;
    static constexpr int TPT = (NT + 127) / 128;        // 4x4 output tiles per thread
    static constexpr int GROUPS = NT >= 128 ? 1 : 128 / NT;
};

#ifdef EMT_PROFILE
// measurement build (-DEMT_PROFILE, scratch/edge_phase_probe.py): cycles per phase of worker group 0 / warp 0 of every CTA
__device__ unsigned long long g_emt_dbg[16];
#define EMT_T(i) do { if (prof) { const long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; } } while (0)
#else
#define EMT_T(i) do { } while (0)
#endif

struct EMTParams {
    const float* ea;
    const int* eperm;
    const float* gout;
    const float *w1, *w2, *w3, *w4;
    int64_t E;
    float* out;       // forward
    float* dea;       // backward, mode 2
    float* partial;   // backward: [gridDim.x][NT * 16]
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ float emt_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

__device__ __forceinline__ void emt_st8(uint32_t ta, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void emt_st16(uint32_t ta, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta),
                 "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
                 "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
                 : "memory");
}
__device__ __forceinline__ void emt_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void emt_ld16(uint32_t ta, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
          "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(ta)
        : "memory");
}
__device__ __forceinline__ void emt_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void emt_mma(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool emt_elect() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void emt_group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// byte offset of element (row n, column c) of a K-major SWIZZLE_128B plane (rows of 32 floats, 8-row groups of 1024 bytes)
__device__ __forceinline__ uint32_t emt_sw128(int n, int c) { return (uint32_t)n * 128u + ((uint32_t)(((c >> 2) ^ (n & 7)) << 4) | (uint32_t)((c & 3) << 2)); }

__device__ __forceinline__ void emt_put(uint8_t* plane, int n, int c, float v, bool lo) {
    const float h = tf32_rn(v);
    *reinterpret_cast<float*>(plane + emt_sw128(n, c)) = lo ? v - h : h;
}

// the group's tile is ready in tensor memory: one elected thread of the group's first warp issues `issue()` and commits to `bar`;
// everybody waits for the MMAs to retire
template <typename F>
__device__ __forceinline__ void emt_round(int barid, int warp_in_wg, uint64_t* bar, uint32_t& ph, F issue) {
    emt_st_wait();
    tc_fence_before();
    emt_group_sync(barid);
    if (warp_in_wg == 0) {
        if (emt_elect()) {
            tc_fence_after();
            issue();
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, ph);
    ph ^= 1;
    tc_fence_after();
}

// MODE 0: forward (out).  MODE 1: backward, weight gradients only.  MODE 2: backward with d ea.
template <int K, int MODE>
__global__ void __launch_bounds__(EMT<K>::nwg(MODE) * 128, 1) k_edge_mlp_tc(const EMTParams p) {
    using T = EMT<K>;
    using C = EMC<K>;
    constexpr int P = T::P, HP = T::HP, TP = T::TP, H = C::H;
    constexpr int NWG = T::nwg(MODE);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // (offset arithmetic keeps the shared address space)
    uint8_t* b1 = smem;
    uint8_t* b2 = b1 + T::B1_BYTES;
    uint8_t* b3 = b2 + T::B2_BYTES;
    uint8_t* b4 = b3 + T::B3_BYTES;
    uint8_t* tail = smem + T::plane_bytes(MODE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);               // [NWG]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 48);
    float* stage_all = reinterpret_cast<float*>(tail + 64);            // [NWG][128][S] (backward only)

    const int wg = threadIdx.x >> 7, t = threadIdx.x & 127, warp_in_wg = t >> 5;

    // ------------------------------------------------------------------ weight planes (hi rows, then lo rows), built by all threads
    for (int i = threadIdx.x; i < 2 * T::N1 * P; i += blockDim.x) {
        const int n = i / P, c = i % P;
        const bool lo = n >= T::N1;
        const int nn = lo ? n - T::N1 : n, b = nn / HP, j = nn % HP;
        const float* w = b == 0 ? p.w1 : (b == 1 ? p.w2 : p.w3);
        emt_put(b1, n, c, (j < H && c < K) ? __ldg(w + j * K + c) : 0.f, lo);
    }
    for (int i = threadIdx.x; i < T::TB * 32 * 32; i += blockDim.x) {
        const int tb = i / 1024, n = (i / 32) % 32, c = i % 32;
        const bool lo = n >= 16;
        const int k = n & 15, q = tb * 32 + c, part = q / HP, j = q % HP;
        emt_put(b2 + tb * 4096, n, c, (k < K && j < H) ? __ldg(p.w4 + k * C::T + part * H + j) : 0.f, lo);
    }
    if constexpr (MODE > 0) {
        for (int i = threadIdx.x; i < 2 * T::N3 * P; i += blockDim.x) {
            const int n = i / P, c = i % P;
            const bool lo = n >= T::N3;
            const int q = lo ? n - T::N3 : n, part = q / HP, j = q % HP;
            emt_put(b3, n, c, (c < K && j < H) ? __ldg(p.w4 + c * C::T + part * H + j) : 0.f, lo);
        }
    }
    if constexpr (MODE > 1) {
        for (int i = threadIdx.x; i < T::DB * 32 * 32; i += blockDim.x) {
            const int db = i / 1024, n = (i / 32) % 32, c = i % 32;
            const bool lo = n >= 16;
            const int ii = n & 15, q = db * 32 + c, b = q / HP, j = q % HP;
            const float* w = b == 0 ? p.w1 : (b == 1 ? p.w2 : p.w3);
            emt_put(b4 + db * 4096, n, c, (ii < K && q < T::DPP && j < H) ? __ldg(w + j * K + ii) : 0.f, lo);
        }
    }
    fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        for (int g = 0; g < NWG; ++g) mbar_init(bars + g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ------------------------------------------------------------------ per-group state
    const uint32_t Xm = tmem_base + (uint32_t)wg * T::slot(MODE), Ym = Xm + T::xc(MODE);     // MMA operand addresses (lane 0)
    const uint32_t lane_sel = (uint32_t)(warp_in_wg * 32) << 16;
    const uint32_t X = Xm + lane_sel, Y = Ym + lane_sel;                                      // this warp's lane quarter
    uint64_t* bar = bars + wg;
    uint32_t ph = 0;
    const int barid = 1 + wg;
    // B descriptors: hi rows first, lo rows behind them (whole 8-row groups: every offset is a multiple of 1024 bytes)
    const uint64_t d1h = make_kmajor_sw128_desc(smem_u32(b1)), d1l = d1h + (uint64_t)((T::N1 * 128) >> 4);
    const uint64_t d2h = make_kmajor_sw128_desc(smem_u32(b2)), d2l = d2h + (uint64_t)((16 * 128) >> 4);
    const uint64_t d3h = make_kmajor_sw128_desc(smem_u32(b3)), d3l = d3h + (uint64_t)((T::N3 * 128) >> 4);
    const uint64_t d4h = make_kmajor_sw128_desc(smem_u32(b4)), d4l = d4h + (uint64_t)((16 * 128) >> 4);
    constexpr uint32_t ID1 = make_idesc_tf32_mn(128, T::N1), ID2 = make_idesc_tf32_mn(128, T::N2), ID3 = make_idesc_tf32_mn(128, T::N3);
    // the three products of one k-step: raw x hi, raw x lo, lo x hi
    auto mma3 = [&](uint32_t d, uint32_t a_raw, uint32_t a_lo, uint64_t bh, uint64_t bl, uint32_t idesc, bool first) {
        emt_mma(d, a_raw, bh, idesc, first ? 0u : 1u);
        emt_mma(d, a_raw, bl, idesc, 1u);
        emt_mma(d, a_lo, bh, idesc, 1u);
    };

    float* stage = stage_all + (size_t)wg * 128 * C::S;
    float* row = stage + (size_t)t * C::S;

    // phase 2: this thread's 4x4 tiles of the weight-gradient matrices  M1 = d_pre4^T tmp,  M2 = d_pre123^T in
    int offA[T::TPT], offB[T::TPT];
    bool accum[T::TPT];
    float acc[T::TPT][4][4];
    const int group = T::TPT == 1 ? t / T::NT : 0;
    if constexpr (MODE > 0) {
#pragma unroll
        for (int u = 0; u < T::TPT; ++u) {
            const int tl = T::TPT == 1 ? t % T::NT : t + 128 * u;
            accum[u] = T::TPT == 1 ? group < T::GROUPS : tl < T::NT;
            if (tl < C::NT1) {
                offA[u] = C::OFF_D4 + 4 * (tl / (C::T / 4));
                offB[u] = C::OFF_TMP + 4 * (tl % (C::T / 4));
            } else {
                const int t2 = (tl < T::NT ? tl : C::NT1) - C::NT1;
                offA[u] = C::OFF_D + 4 * (t2 / (C::KP / 4));
                offB[u] = C::OFF_IN + 4 * (t2 % (C::KP / 4));
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[u][r][c] = 0.f;
        }
    }

    const int64_t ntiles = (p.E + 127) / 128;
    const int64_t tstride = (int64_t)gridDim.x * NWG;
    // this thread's row of ea is fetched one tile ahead (the upstream gradient row is loaded at the top of its own tile and
    // first used two round trips later)
    float in_n[C::KP];
    int64_t src_n = 0;
    auto fetch = [&](int64_t tile) {
        const int64_t e = tile * 128 + t;
#pragma unroll
        for (int i = 0; i < C::KP; ++i) in_n[i] = 0.f;
        src_n = 0;
        if (tile < ntiles && e < p.E) {
            src_n = p.eperm ? (int64_t)__ldg(p.eperm + e) : e;
            load_edge_row<K>(p.ea + src_n * K, in_n);
        }
    };
    fetch((int64_t)blockIdx.x * NWG + wg);
#ifdef EMT_PROFILE
    const bool prof = t == 0;
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
    const long long tbegin = tlast;
#endif
    for (int64_t tile = (int64_t)blockIdx.x * NWG + wg; tile < ntiles; tile += tstride) {
        const int64_t e = tile * 128 + t;
        const bool live = e < p.E;
        const int64_t src = src_n;
        float in[P];
#pragma unroll
        for (int i = 0; i < P; ++i) in[i] = i < C::KP ? in_n[i < C::KP ? i : 0] : 0.f;
        // ---------------------------------------------------------------- layer 1: pre123 = ea [W1;W2;W3]^T -> X [0, 6P)
        {
            float lo[P];
#pragma unroll
            for (int i = 0; i < P; ++i) lo[i] = emt_lo(in[i]);
            if constexpr (P == 8) {
                emt_st8(Y, in);
                emt_st8(Y + 8, lo);
            } else {
                emt_st16(Y, in);
                emt_st16(Y + 16, lo);
            }
        }
        if constexpr (MODE > 0) {
#pragma unroll
            for (int i = 0; i < C::KP; i += 4) *reinterpret_cast<float4*>(row + C::OFF_IN + i) = make_float4(in[i], in[i + 1], in[i + 2], in[i + 3]);
        }
        // Global loads are issued only now, AFTER the last use of the previous fetch: the hardware scoreboard counts per slot,
        // not per register, so a consumer of the old row that waits on its slot would also wait for loads issued before it
        // (ncu: 25 % of the forward's stall samples sat on the first emt_lo with the fetch placed at the top of the tile).
        float go[C::KP];
#pragma unroll
        for (int i = 0; i < C::KP; ++i) go[i] = 0.f;
        if constexpr (MODE > 0) {
            if (live) load_edge_row<K>(p.gout + e * K, go);
        }
        fetch(tile + tstride);
        EMT_T(0);      // loads + first stores
        emt_round(barid, warp_in_wg, bar, ph, [&] {
#pragma unroll
            for (int s = 0; s < P / 8; ++s) mma3(Xm, Ym + 8 * s, Ym + P + 8 * s, d1h + (uint64_t)(2 * s), d1l + (uint64_t)(2 * s), ID1, s == 0);
        });
        EMT_T(1);      // round trip 1
        // ---------------------------------------------------------------- activations, 16 hidden units at a time, in place:
        // reads   pre1 [c16, +16) | pre2 [2P + c16) | pre3 [4P + c16)
        // writes  relu raw [c16) | product raw [2P + c16) | relu lo [4P + c16) | product lo [6P + c16)   (= A operand of layer 2:
        //         raw tmp in columns [0, 4P), residuals in [4P, 8P))
        uint32_t mask1 = 0;
#pragma unroll
        for (int c = 0; c < HP / 16; ++c) {
            float r[16], a2[16], a3[16];
            emt_ld16(X + 16 * c, r);
            emt_ld16(X + 2 * P + 16 * c, a2);
            emt_ld16(X + 4 * P + 16 * c, a3);
            emt_ld_wait();
            float pr[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE > 0 && r[i] > 0.f) mask1 |= 1u << (16 * c + i);
                r[i] = fmaxf(r[i], 0.f);
                a2[i] = tanh_fast5(a2[i]);
                a3[i] = tanh_fast5(a3[i]);
                pr[i] = a2[i] * a3[i];
            }
            emt_st16(X + 16 * c, r);
            emt_st16(X + 2 * P + 16 * c, pr);
            if constexpr (MODE > 0) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const int j0 = 16 * c + i;
                    if (j0 < H) {
                        *reinterpret_cast<float4*>(row + C::OFF_TMP + j0) = make_float4(r[i], r[i + 1], r[i + 2], r[i + 3]);
                        *reinterpret_cast<float4*>(row + C::OFF_TMP + H + j0) = make_float4(pr[i], pr[i + 1], pr[i + 2], pr[i + 3]);
                        // the two tanh factors wait in the d_pre2 / d_pre3 slots until the upstream gradient is known
                        *reinterpret_cast<float4*>(row + C::OFF_D + H + j0) = make_float4(a2[i], a2[i + 1], a2[i + 2], a2[i + 3]);
                        *reinterpret_cast<float4*>(row + C::OFF_D + 2 * H + j0) = make_float4(a3[i], a3[i + 1], a3[i + 2], a3[i + 3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                r[i] = emt_lo(r[i]);
                pr[i] = emt_lo(pr[i]);
            }
            emt_st16(X + 4 * P + 16 * c, r);
            emt_st16(X + 6 * P + 16 * c, pr);
        }
        EMT_T(2);      // activations
        // ---------------------------------------------------------------- layer 2: pre4 = tmp W4^T -> Y [0, 16)
        emt_round(barid, warp_in_wg, bar, ph, [&] {
#pragma unroll
            for (int s = 0; s < TP / 8; ++s) {
                const uint64_t adv = (uint64_t)((s / 4) * (4096 >> 4) + 2 * (s % 4));
                mma3(Ym, Xm + 8 * s, Xm + TP + 8 * s, d2h + adv, d2l + adv, ID2, s == 0);
            }
        });
        EMT_T(3);      // round trip 2
        float pre4[16];
        emt_ld16(Y, pre4);
        emt_ld_wait();
        if constexpr (MODE == 0) {
            if (live) {
                float* op = p.out + e * K;
#pragma unroll
                for (int k = 0; k < K; k += 2) *reinterpret_cast<float2*>(op + k) = make_float2(fmaxf(pre4[k], 0.f), fmaxf(pre4[k + 1], 0.f));
            }
        } else {
            // ------------------------------------------------------------ d pre4 = gout * relu'(pre4);  d tmp = d pre4 W4 -> X [0, 4P)
            float dp4[P];
#pragma unroll
            for (int k = 0; k < P; ++k) dp4[k] = (k < K && pre4[k] > 0.f) ? go[k < C::KP ? k : 0] : 0.f;
#pragma unroll
            for (int k = 0; k < C::KP; k += 4) *reinterpret_cast<float4*>(row + C::OFF_D4 + k) = make_float4(dp4[k], dp4[k + 1], dp4[k + 2], dp4[k + 3]);
            {
                float lo[P];
#pragma unroll
                for (int i = 0; i < P; ++i) lo[i] = emt_lo(dp4[i]);
                if constexpr (P == 8) {
                    emt_st8(Y, dp4);
                    emt_st8(Y + 8, lo);
                } else {
                    emt_st16(Y, dp4);
                    emt_st16(Y + 16, lo);
                }
            }
            emt_round(barid, warp_in_wg, bar, ph, [&] {
#pragma unroll
                for (int s = 0; s < P / 8; ++s) mma3(Xm, Ym + 8 * s, Ym + P + 8 * s, d3h + (uint64_t)(2 * s), d3l + (uint64_t)(2 * s), ID3, s == 0);
            });
            EMT_T(4);  // d pre4 + round trip 3
            // d tmp: relu part in X [c16), product part in X [2P + c16).  In place (mode 2):
            // d pre1 raw [c16) | d pre2 raw [2P + c16) | d pre3 raw [4P + c16) | residuals 6P further
#pragma unroll
            for (int c = 0; c < HP / 16; ++c) {
                float dt1[16], dt2[16];
                emt_ld16(X + 16 * c, dt1);
                emt_ld16(X + 2 * P + 16 * c, dt2);
                emt_ld_wait();
                float g1[16], g2[16], g3[16];
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const int j0 = 16 * c + i;
                    float4 t2 = make_float4(0.f, 0.f, 0.f, 0.f), t3 = t2;
                    if (j0 < H) {
                        t2 = *reinterpret_cast<const float4*>(row + C::OFF_D + H + j0);
                        t3 = *reinterpret_cast<const float4*>(row + C::OFF_D + 2 * H + j0);
                    }
                    const float t2a[4] = {t2.x, t2.y, t2.z, t2.w}, t3a[4] = {t3.x, t3.y, t3.z, t3.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        g1[i + jj] = ((mask1 >> (j0 + jj)) & 1u) ? dt1[i + jj] : 0.f;
                        g2[i + jj] = dt2[i + jj] * t3a[jj] * (1.f - t2a[jj] * t2a[jj]);
                        g3[i + jj] = dt2[i + jj] * t2a[jj] * (1.f - t3a[jj] * t3a[jj]);
                    }
                    if (j0 < H) {
                        *reinterpret_cast<float4*>(row + C::OFF_D + j0) = make_float4(g1[i], g1[i + 1], g1[i + 2], g1[i + 3]);
                        *reinterpret_cast<float4*>(row + C::OFF_D + H + j0) = make_float4(g2[i], g2[i + 1], g2[i + 2], g2[i + 3]);
                        *reinterpret_cast<float4*>(row + C::OFF_D + 2 * H + j0) = make_float4(g3[i], g3[i + 1], g3[i + 2], g3[i + 3]);
                    }
                }
                if constexpr (MODE > 1) {
                    emt_st16(X + 16 * c, g1);
                    emt_st16(X + 2 * P + 16 * c, g2);
                    emt_st16(X + 4 * P + 16 * c, g3);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        g1[i] = emt_lo(g1[i]);
                        g2[i] = emt_lo(g2[i]);
                        g3[i] = emt_lo(g3[i]);
                    }
                    emt_st16(X + 6 * P + 16 * c, g1);
                    emt_st16(X + 8 * P + 16 * c, g2);
                    emt_st16(X + 10 * P + 16 * c, g3);
                }
            }
            if constexpr (C::DP > C::D) {
#pragma unroll
                for (int i = C::D; i < C::DP; ++i) row[C::OFF_D + i] = 0.f;
            }
            if constexpr (MODE > 1) {
                // -------------------------------------------------------- d ea = d pre123 [W1;W2;W3] -> Y [0, 16)
                emt_round(barid, warp_in_wg, bar, ph, [&] {
#pragma unroll
                    for (int s = 0; s < T::DPP / 8; ++s) {
                        const uint64_t adv = (uint64_t)((s / 4) * (4096 >> 4) + 2 * (s % 4));
                        mma3(Ym, Xm + 8 * s, Xm + T::DPP + 8 * s, d4h + adv, d4l + adv, ID2, s == 0);
                    }
                });
                float din[16];
                emt_ld16(Y, din);
                emt_ld_wait();
                if (live) {
                    float* dp = p.dea + src * K;
#pragma unroll
                    for (int i = 0; i < K; i += 2) *reinterpret_cast<float2*>(dp + i) = make_float2(din[i], din[i + 1]);
                }
            }
            EMT_T(5);  // pre-activation gradients (+ d ea round trip)
            // ------------------------------------------------------------ phase 2: outer products over the group's 128 staged edges
            emt_group_sync(barid);
            EMT_T(6);  // group barrier before phase 2
#pragma unroll
            for (int u = 0; u < T::TPT; ++u) {
                if (accum[u]) {
                    const float* pa = stage + (size_t)group * C::S + offA[u];
                    const float* pb = stage + (size_t)group * C::S + offB[u];
#pragma unroll 4
                    for (int r = group; r < 128; r += T::GROUPS) {
                        const float4 a = *reinterpret_cast<const float4*>(pa);
                        const float4 b = *reinterpret_cast<const float4*>(pb);
                        pa += T::GROUPS * C::S;
                        pb += T::GROUPS * C::S;
                        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[u][i][j] = fmaf(av[i], bv[j], acc[u][i][j]);
                    }
                }
            }
            emt_group_sync(barid);
            EMT_T(7);  // phase 2
        }
        EMT_T(8);      // output store
    }
#ifdef EMT_PROFILE
    if (prof && wg == 0) {
        for (int i = 0; i < 9; ++i) atomicAdd(&g_emt_dbg[i], (unsigned long long)tacc[i]);
        atomicAdd(&g_emt_dbg[9], (unsigned long long)(clock64() - tbegin));
        atomicAdd(&g_emt_dbg[10], 1ull);
    }
#endif
    if constexpr (MODE > 0) {
        // reduce the thread groups of this worker group in a fixed order and emit its partial (tile layout of k_edge_mlp_bwd)
        float* red = stage;   // [GROUPS][NT * 16]
#pragma unroll
        for (int u = 0; u < T::TPT; ++u) {
            if (accum[u]) {
                const int tl = T::TPT == 1 ? t % T::NT : t + 128 * u;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) red[(size_t)group * T::NT * 16 + tl * 16 + i * 4 + j] = acc[u][i][j];
            }
        }
        emt_group_sync(barid);
        for (int i = t; i < T::NT * 16; i += 128) {
            float s = 0.f;
#pragma unroll
            for (int g = 0; g < T::GROUPS; ++g) s += red[(size_t)g * T::NT * 16 + i];
            red[i] = s;                                   // (index i of every thread group is read and written by this thread only)
        }
        // ... then the worker groups of the CTA, again in a fixed order: ONE partial per CTA for k_edge_mlp_bwd_reduce
        __syncthreads();
        float* dst = p.partial + (size_t)blockIdx.x * T::NT * 16;
        for (int i = threadIdx.x; i < T::NT * 16; i += NWG * 128) {
            float s = 0.f;
#pragma unroll
            for (int g = 0; g < NWG; ++g) s += stage_all[(size_t)g * 128 * C::S + i];
            dst[i] = s;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

template <int K, int MODE>
static int emt_launch(const EMTParams& p, int* nparts, cudaStream_t st) {
    using T = EMT<K>;
    static bool configured[64] = {};
    if (auto once_ = first_use_on_device(configured))
        GNNML3_CUDA(cudaFuncSetAttribute(k_edge_mlp_tc<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::smem_bytes(MODE)));
    constexpr int NWG = T::nwg(MODE);
    const int64_t ntiles = (p.E + 127) / 128;
    int64_t grid = (ntiles + NWG - 1) / NWG;
    if (grid > kNumSMs) grid = kNumSMs;
    k_edge_mlp_tc<K, MODE><<<(int)grid, NWG * 128, T::smem_bytes(MODE), st>>>(p);
    GNNML3_LAUNCH_CHECK();
    if (nparts) *nparts = (int)grid;
    return GNNML3_OK;
}

#define EMT_DISPATCH(KV, ...)                                      \
    switch (KV) {                                                  \
        case 2: { constexpr int K_ = 2; __VA_ARGS__; } break;      \
        case 4: { constexpr int K_ = 4; __VA_ARGS__; } break;      \
        case 6: { constexpr int K_ = 6; __VA_ARGS__; } break;      \
        case 8: { constexpr int K_ = 8; __VA_ARGS__; } break;      \
        case 10: { constexpr int K_ = 10; __VA_ARGS__; } break;    \
        case 12: { constexpr int K_ = 12; __VA_ARGS__; } break;    \
        case 14: { constexpr int K_ = 14; __VA_ARGS__; } break;    \
        case 16: { constexpr int K_ = 16; __VA_ARGS__; } break;    \
        default: return set_err(GNNML3_ERR_INVALID, "edge_mlp_tc: K=%d not instantiated (even K <= 16)", KV); \
    }

// called by gnnml3_edge_mlp_fwd / _bwd (edge_mlp.cu) after argument validation
int edge_mlp_tc_fwd(const float* ea, const int32_t* eperm, const float* w1, const float* w2, const float* w3, const float* w4, int64_t E,
                    int K, float* out, cudaStream_t st) {
    EMTParams p;
    p.ea = ea; p.eperm = eperm; p.gout = nullptr; p.w1 = w1; p.w2 = w2; p.w3 = w3; p.w4 = w4; p.E = E; p.out = out; p.dea = nullptr;
    p.partial = nullptr;
    EMT_DISPATCH(K, return (emt_launch<K_, 0>(p, nullptr, st)));
    return GNNML3_OK;
}

size_t edge_mlp_tc_bwd_workspace_bytes(int K) {
    const int KP = pad4(K), T = 4 * K, DP = pad4(6 * K);
    const size_t nt = (size_t)(KP / 4) * (T / 4) + (size_t)(DP / 4) * (KP / 4);
    return (size_t)kNumSMs * nt * 16 * sizeof(float);
}

int edge_mlp_tc_bwd(const float* ea, const int32_t* eperm, const float* gout, const float* w1, const float* w2, const float* w3,
                    const float* w4, int64_t E, int K, float* dea, float* dw1, float* dw2, float* dw3, float* dw4, float* partial,
                    cudaStream_t st) {
    EMTParams p;
    p.ea = ea; p.eperm = eperm; p.gout = gout; p.w1 = w1; p.w2 = w2; p.w3 = w3; p.w4 = w4; p.E = E; p.out = nullptr; p.dea = dea;
    p.partial = partial;
    int nparts = 0, rc;
    if (dea) {
        EMT_DISPATCH(K, rc = (emt_launch<K_, 2>(p, &nparts, st)));
    } else {
        EMT_DISPATCH(K, rc = (emt_launch<K_, 1>(p, &nparts, st)));
    }
    if (rc) return rc;
    EMT_DISPATCH(K, (k_edge_mlp_bwd_reduce<K_><<<cdiv(EMC<K_>::NOUT, 32), 256, 0, st>>>(partial, nparts, dw1, dw2, dw3, dw4)));
    GNNML3_LAUNCH_CHECK();
    return GNNML3_OK;
}

}  // namespace gnnml3

#ifdef EMT_PROFILE
extern "C" __attribute__((visibility("default"))) int gnnml3_emt_debug_fetch(unsigned long long* host16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host16, gnnml3::g_emt_dbg, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {};
        cudaMemcpyToSymbol(gnnml3::g_emt_dbg, z, sizeof(z));
    }
    return 0;
}
#endif
