/*
 * ifl_sweep2.cuh -- applyPreconditioner F3:495 (unmasked), TWO ROWS PER LANE.  Included by ifl_solve_kernels.cu.
 *
 * k_precon_sweep (one row per lane) spends most of a sweep in the skew of its strips: lane l runs l macro-steps behind
 * lane 0, so a strip hands its last row to the next strip 31 macro-steps after it started, and 31 steps x nstrips is
 * the larger part of the (w/SKB + h) step critical path.  Here a lane owns the 2 x SKB tile (rows 2L and 2L+1, SKB
 * columns) of its column block: 16 lanes cover the 32 rows of a strip, the skew is 15 macro-steps, and the two rows of
 * a lane pipeline inside the lane (row 2L+1 needs row 2L of the SAME macro-step: registers, no exchange).  Per cell the
 * arithmetic is the reference's, operation for operation (F3:500-506 / F3:514-520).
 *
 * Data layout in HBM is unchanged (strip-skewed vectors, packed coefficients): cell (c, l) -- column block c, row l of
 * the strip -- lives in m-row c + l.  At macro-step T lane L handles column block c = T - L, i.e. m-rows T + L (row 2L)
 * and T + L + 1 (row 2L+1): the lanes of a warp read 17 consecutive m-rows, which the loader warp keeps in a circular
 * buffer per operand array ([array][S*U m-rows][32][SKB]; tile q at position q mod S).  T runs 0 .. NT-1 in the forward
 * sweep and NT-1 .. 0 in the backward sweep (NT = ceil8(cw + 15)); everything else mirrors.
 *
 * Hand-over between strips, tickets, credits, flag-in-data rings and boundary rows are those of k_precon_sweep with
 * the skew depth 31 replaced by 15: a producer publishes from its macro-step 8 on, consumer step c uses producer step
 * c + 15.  Cells outside the grid are zero padding inside the strip's own stream (tl >= cw + 39), they produce +0.0 and
 * storing +0.0 keeps them zero, so no block needs per-cell predicates; only stores into a global boundary row (w
 * entries, no padding) are predicated.
 */
#pragma once

static constexpr int SW2_D = 15;                                  /* skew depth: lane L is L macro-steps behind lane 0 */
static constexpr int SW2_ROWB = 32 * SKB * (int)sizeof(double);    /* bytes per m-row */
static constexpr int SW2_TILEB = SW_U * SW2_ROWB;                  /* bytes per tile and array */
static constexpr int SW2_ARRB = SW_STAGES * SW2_TILEB;             /* bytes per operand array ring */
static constexpr int SW2_ROWS = SW_STAGES * SW_U;                  /* m-rows per ring */

/* operands of one macro-step of one lane: [row][cell] */
template <bool WITH_DOT>
struct CellOps2 {
    double a[2][SKB], cx[2][SKB], cy[2][SKB], pr[2][SKB];
};
/* pa / pb: shared addresses of this lane's cells (row 2L position) in m-rows m and m + 1 of the `src` ring */
__device__ __forceinline__ void load_step2(double (&a)[2][SKB], double (&cx)[2][SKB], double (&cy)[2][SKB],
                                           double (&pr)[2][SKB], unsigned pa, unsigned pb) {
    lds_cells(pa, a[0]);
    lds_cells(pa + SW2_ARRB, cx[0]);
    lds_cells(pa + 2 * SW2_ARRB, cy[0]);
    lds_cells(pa + 3 * SW2_ARRB, pr[0]);
    lds_cells(pb + 16u, a[1]);
    lds_cells(pb + 16u + SW2_ARRB, cx[1]);
    lds_cells(pb + 16u + 2 * SW2_ARRB, cy[1]);
    lds_cells(pb + 16u + 3 * SW2_ARRB, pr[1]);
}

/* Exchange store: the lane at the outgoing edge of a ring-out strip stores through the generic (mapped, possibly remote)
 * address of the next strip's ring slot; every other lane stores into its own exchange slot with a plain shared-memory
 * store, so that the shared-memory load of the adjacent lane that follows is not ordered behind a generic store. */
__device__ __forceinline__ void xchg_store2(bool to_ring, double* ring_slot, unsigned own_slot, double v) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q st.f64 [%1], %3;\n@!q st.shared.f64 [%2], %3;\n}"
                 ::"r"((unsigned)to_ring), "l"(ring_slot), "r"(own_slot), "d"(v) : "memory");
}
/* predicated store into a global boundary row (state space given: no generic address resolution) */
__device__ __forceinline__ void st_global_if(bool p, double* addr, double v) {
    asm volatile("{\n.reg .pred q;\n.reg .u64 ga;\nsetp.ne.u32 q, %0, 0;\ncvta.to.global.u64 ga, %1;\n@q st.global.f64 [ga], %2;\n}"
                 ::"r"((unsigned)p), "l"(addr), "d"(v) : "memory");
}

template <int DIR, bool WITH_DOT>
__global__ void __launch_bounds__(2 * SW_WARPS * 32)
k_precon_sweep2(SkewGeom g, const double* src, double* dst, const double* __restrict__ pack,
                const double* __restrict__ rr, double* bnd, double* bnd_peer, double* strip_part,
                SolveScalars* sc, int dot_mode, int check_done, long long* trace) {
    static_assert(SKB == 2 && SW_U == 8 && SW_FG == 8, "two-row sweep: 2 x 2 tiles, 8-step blocks, 8-row staged tiles");
    namespace cg = cooperative_groups;
    constexpr int U = SW_U, UB = SW_FG, S = SW_STAGES, D = SW2_D;
    constexpr int NARR = WITH_DOT ? 5 : 4;
    constexpr int E_IN = DIR > 0 ? 0 : 15;
    constexpr int E_OUT = DIR > 0 ? 15 : 0;
    extern __shared__ __align__(128) unsigned char wf_smem[];
    __shared__ __align__(16) double s_inbox[SW_WARPS][SW_RB * SKB];
    __shared__ volatile int s_credit[SW_WARPS];
    __shared__ __align__(16) double s_scratch[SW_WARPS][SW_FG * SKB];
    constexpr int XS = SW_FG * SKB + 1;
    __shared__ __align__(16) double s_xchg[SW_WARPS][32 * XS];
    __shared__ __align__(16) double s_zero[SW_FG * SKB];
    __shared__ int s_ticket;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    if (check_done && sc->done) return;
    const long long tr_enter = trace ? (long long)globaltimer_ns() : 0;
    const int lane = threadIdx.x & 31;
    const bool loader = threadIdx.x >= SW_WARPS * 32;
    const int warp = (threadIdx.x >> 5) - (loader ? SW_WARPS : 0);
    /* per strip: NARR operand rings, then S full + S empty mbarriers */
    unsigned char* ring = wf_smem + (size_t)warp * S * (((WITH_DOT ? 2 : 1) * SW_TILE + SW_PACK) * sizeof(double));
    static_assert(((WITH_DOT ? 2 : 1) * SW_TILE + SW_PACK) * sizeof(double) * SW_STAGES == (size_t)NARR * SW2_ARRB, "ring size");
    unsigned long long* bar_full = reinterpret_cast<unsigned long long*>(
        wf_smem + (size_t)SW_WARPS * NARR * SW2_ARRB) + warp * 2 * S;
    unsigned long long* bar_empty = bar_full + S;
    if (!loader) {
        for (int i = lane; i < 32 * XS; i += 32) s_xchg[warp][i] = 0.0;
        if (warp == 0 && lane < SW_FG * SKB) s_zero[lane] = 0.0;
        for (int i = lane; i < SW_RB * SKB; i += 32) s_inbox[warp][i] = sentinel();
        if (lane == 0) {
            s_credit[warp] = 0;
            for (int i = 0; i < S; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    if (crank == 0 && threadIdx.x == 0) {
        int tk = (int)atomicAdd(DIR > 0 ? &sc->ticket_f : &sc->ticket_b, 1u);
        for (int r = 0; r < csize; r++) *cluster.map_shared_rank(&s_ticket, r) = tk;
    }
    cluster.sync();
    const int nown = g.s1 - g.s0;
    const int cslot = crank * SW_WARPS + warp;
    const int slot = s_ticket * csize * SW_WARPS + cslot;
    const bool live = slot < nown;
    const int k = DIR > 0 ? g.s0 + slot : g.s1 - 1 - slot;
    const int w = g.w;
    const int cw = (w + SKB - 1) / SKB;
    const int nblk = (cw + D + UB - 1) / UB;                          /* blocks of UB macro-steps */
    const int ntile = nblk + 2;                                       /* tiles 0 .. nblk+1 hold the m-rows 0 .. 8*nblk + 15 */
    double acc = 0.0;
    if (live && loader) {
        if (lane == 0) {
            /* tile ordinal j (the order of use): tile q = j (forward) / ntile-1-j (backward), ring position q mod S.  One
             * ordinal more than the strip has: the strip warp prefetches across the last block without a test. */
            const double* gsrc0 = src + (long long)k * g.ss;
            const double* grr0 = WITH_DOT ? rr + (long long)k * g.ss : nullptr;
            const double* gpack0 = pack + (long long)k * (g.tl / U) * SW_PACK;
            const unsigned ring0 = smem_u32(ring), full0 = smem_u32(bar_full), empty0 = smem_u32(bar_empty);
            int q = DIR > 0 ? 0 : ntile - 1;
            int pos = q % S;
            int jm = 0;                                               /* j mod S */
            unsigned lap = 0;                                         /* j / S */
            for (int j = 0; j <= ntile; j++) {
                if (lap > 0) mbar_wait_u32(empty0 + 8u * pos, (lap - 1u) & 1u);
                const int qa = q < 0 ? 0 : (q >= ntile ? ntile - 1 : q);      /* the extra ordinal re-reads an existing tile */
                const unsigned fb = full0 + 8u * pos;
                const unsigned sdst = ring0 + (unsigned)(pos * SW2_TILEB);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((unsigned)(NARR * SW2_TILEB)) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sdst), "l"(gsrc0 + (long long)qa * SW_TILE), "r"((unsigned)SW2_TILEB), "r"(fb) : "memory");
#pragma unroll
                for (int a = 0; a < 3; a++)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(sdst + (unsigned)((a + 1) * SW2_ARRB)), "l"(gpack0 + (long long)qa * SW_PACK + a * SW_TILE),
                                   "r"((unsigned)SW2_TILEB), "r"(fb) : "memory");
                if (WITH_DOT)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(sdst + (unsigned)(4 * SW2_ARRB)), "l"(grr0 + (long long)qa * SW_TILE), "r"((unsigned)SW2_TILEB), "r"(fb) : "memory");
                q += DIR;
                pos += DIR;
                if (pos == S) pos = 0;
                if (pos < 0) pos = S - 1;
                if (++jm == S) { jm = 0; lap++; }
            }
        }
    } else if (live && lane < 16) {
        constexpr unsigned HALF = 0x0000ffffu;                        /* the 16 lanes of a strip */
        const int tfirst = DIR > 0 ? 0 : nblk * UB - 1;
        const bool has_prev = DIR > 0 ? (k > 0) : (k < g.nstrips - 1);
        const bool is_out = lane == E_OUT;
        const bool is_in = lane == E_IN;
        const bool lane0 = lane == 0;

        const bool in_ring = cslot > 0;
        EdgeSrc feed;
        feed.mode = !has_prev ? 0 : (in_ring ? 2 : 1);
        feed.src = in_ring ? &s_inbox[warp][0] : (has_prev ? bnd + (long long)(k - DIR) * w : nullptr);
        feed.ring = in_ring ? &s_inbox[warp][0] : nullptr;
        feed.w = w;
        feed.feeder = is_in && feed.mode != 0;
        const bool use_feed = feed.mode != 0;
        const bool ring_feed = feed.mode == 2;
        const bool out_ring = (cslot < csize * SW_WARPS - 1) && (slot + 1 < nown);
        const bool to_peer = bnd_peer != nullptr && (DIR > 0 ? k == g.s1 - 1 : k == g.s0);
        double* out_base;
        volatile int* credit_of_prev = nullptr;
        if (out_ring) out_base = cluster.map_shared_rank(&s_inbox[(cslot + 1) % SW_WARPS][0], (cslot + 1) / SW_WARPS);
        else out_base = (to_peer ? bnd_peer : bnd) + (long long)k * w;
        if (in_ring) credit_of_prev = cluster.map_shared_rank((int*)&s_credit[(cslot - 1) % SW_WARPS], (cslot - 1) / SW_WARPS);
        const bool pub_credit = lane0 && credit_of_prev != nullptr;
        double* const oscratch = &s_scratch[warp][0];
        const unsigned xw_own32 = pin_u32(smem_u32(&s_xchg[warp][lane * XS]));
        const unsigned xr_nb = pin_u32(smem_u32(&s_xchg[warp][((lane - DIR) & 31) * XS]));
        const unsigned xr_zero = pin_u32(smem_u32(&s_zero[0]));

        /* row 2L of column block c = T - L sits in m-row T + L at lane position 2L; row 2L+1 in m-row T + L + 1 at 2L + 1 */
        double* pdst = dst + (long long)k * g.ss + ((long long)(tfirst + lane) * 32 + 2 * lane) * SKB;
        double dprev[2][SKB];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int j = 0; j < SKB; j++) dprev[r][j] = 0.0;
        double cxprev[2] = {0.0, 0.0};
        long long tr_first = 0;

        const unsigned full0 = pin_u32(smem_u32(bar_full));
        const unsigned empty0 = pin_u32(smem_u32(bar_empty));
        const unsigned inbox0 = pin_u32(smem_u32(&s_inbox[warp][0]));
        /* operand cursor: pa / pb = this lane's position in the m-rows m and m + 1 that the NEXT operand fetch reads */
        const unsigned rbase = pin_u32(smem_u32(ring) + (unsigned)(2 * lane * SKB * (int)sizeof(double)));
        const unsigned rend = rbase + (unsigned)SW2_ARRB;
        unsigned pa = rbase + (unsigned)(((tfirst + lane) % SW2_ROWS) * SW2_ROWB);
        unsigned pb = rbase + (unsigned)(((tfirst + lane + 1) % SW2_ROWS) * SW2_ROWB);
        auto advance = [&]() {
            if (DIR > 0) { pa = pb; pb += (unsigned)SW2_ROWB; pb = pb == rend ? rbase : pb; }
            else { pb = pa; pa = (pa == rbase ? rend : pa) - (unsigned)SW2_ROWB; }
        };
        /* tile ordinals: acquire (wait full) runs 3 ahead of release (arrive empty); position = tile index mod S */
        const int pos0 = DIR > 0 ? 0 : (ntile - 1) % S;
        auto pos_step = [&](int p) { p += DIR; return p == S ? 0 : (p < 0 ? S - 1 : p); };
        int acq_pos = pos0, acq_jm = 0; unsigned acq_par = 0;
        int rel_pos = pos0;
        auto acquire = [&]() {
            mbar_wait_u32(full0 + 8u * acq_pos, acq_par);
            acq_pos = pos_step(acq_pos);
            if (++acq_jm == S) { acq_jm = 0; acq_par ^= 1u; }
        };

        CellOps2<WITH_DOT> ops[2];
        acquire(); acquire(); acquire();                              /* ordinals 0, 1, 2: the 17 m-rows of the first macro-step */
        load_step2(ops[0].a, ops[0].cx, ops[0].cy, ops[0].pr, pa, pb);
        unsigned qa = pa, qb = pb;                                    /* where the CURRENT macro-step's rr cells are (WITH_DOT) */
        advance();

        double dn[SKB], epk[2][SKB];
#pragma unroll
        for (int j = 0; j < SKB; j++) { dn[j] = 0.0; epk[0][j] = 0.0; epk[1][j] = 0.0; }
        const unsigned sent_hi = pin_u32(0xFFF8B200u);

        /* One block of UB macro-steps, unrolled.  RINGFEED: incoming edge values come slot by slot from this strip's
         * ring (the lane at the incoming edge loads the slot instead of a neighbour's exchange slot); otherwise they are
         * in ecur (global boundary row) or absent (zeros).  GOUT: the outgoing edge row goes to a global boundary row
         * (predicated stores at po_lane, column block cout0 + DIR*s); otherwise the edge lane's exchange stores already
         * went into the next strip's ring (xw_b). */
        auto run_block = [&](auto ringfeed_tag, auto gout_tag, const int t0, const int sc0, const double (&ecur)[SW_GV],
                             double* po_lane, const bool to_ring) {
            constexpr bool RINGFEED = decltype(ringfeed_tag)::value;
            constexpr bool GOUT = decltype(gout_tag)::value;
            const unsigned in_rest = inbox0 + (unsigned)(((sc0 + D + 1) & (SW_RB - 1)) * SKB * (int)sizeof(double));
            const unsigned in_next = inbox0 + (unsigned)(((sc0 + UB + D + 1) & (SW_RB - 1)) * SKB * (int)sizeof(double));
            const bool is_in_feed = is_in && use_feed && !RINGFEED;
            const unsigned xr_b = is_in ? (RINGFEED ? in_rest : xr_zero) : xr_nb;
            if (!RINGFEED) {
#pragma unroll
                for (int j = 0; j < SKB; j++) dn[j] = is_in_feed ? ecur[j] : dn[j];
            }
#pragma unroll
            for (int s = 0; s < UB; s++) {
                if (RINGFEED) {
                    /* the slot of macro-step s+1 was requested one macro-step ago; it must have arrived before the edge
                     * lane loads it at the end of this macro-step; then request the one after */
                    const unsigned a = in_rest + (unsigned)(s * SKB * (int)sizeof(double));
                    ring_wait(a, epk[(s + 1) & 1], true, true);
                    ring_peek(s + 1 < UB ? a + (unsigned)(SKB * sizeof(double)) : in_next, epk[s & 1]);
                }
                if (s == UB - 1) acquire();                           /* the next macro-step's operands reach into tile ordinal b + 3 */
                CellOps2<WITH_DOT>& nx = ops[(s + 1) & 1];
                const CellOps2<WITH_DOT>& op = ops[s & 1];
                load_step2(nx.a, nx.cx, nx.cy, nx.pr, pa, pb);
                double rr0[SKB], rr1[SKB];
                if (WITH_DOT) { lds_cells(qa + 4 * SW2_ARRB, rr0); lds_cells(qb + 16u + 4 * SW2_ARRB, rr1); }
                qa = pa; qb = pb;
                advance();

                double d0[SKB], d1[SKB], dnn[SKB];
                double* const xg = po_lane + s * SKB;                 /* ring slot (only used by the edge lane of a ring-out strip) */
                const unsigned xo = xw_own32 + (unsigned)(s * SKB * (int)sizeof(double));
                const unsigned xr = xr_b + (unsigned)(s * SKB * (int)sizeof(double));
                if (DIR > 0) {
                    double dl = dprev[0][SKB - 1];
#pragma unroll
                    for (int j = 0; j < SKB; j++) {                   /* row 2L: upper neighbours from the lane above (dn) */
                        double v = op.a[0][j];
                        v = v - cxprev[0] * dl;                       /* (aPlusX[i-1]*precon[i-1]) * dst[i-1]   F3:500 */
                        v = v - op.cy[0][j] * dn[j];                  /* (aPlusY[i-w]*precon[i-w]) * dst[i-w]   F3:503 */
                        cxprev[0] = op.cx[0][j];
                        d0[j] = v * op.pr[0][j];
                        dl = d0[j];
                    }
                    dl = dprev[1][SKB - 1];
#pragma unroll
                    for (int j = 0; j < SKB; j++) {                   /* row 2L+1: upper neighbours are this macro-step's row 2L */
                        double v = op.a[1][j];
                        v = v - cxprev[1] * dl;
                        v = v - op.cy[1][j] * d0[j];
                        cxprev[1] = op.cx[1][j];
                        d1[j] = v * op.pr[1][j];
                        dl = d1[j];
                        xchg_store2(to_ring, xg + j, xo + (unsigned)(j * sizeof(double)), d1[j]);   /* hand the lower row on right away */
                        dnn[j] = xchg_load(xr + (unsigned)(j * sizeof(double)));
                    }
                } else {
                    double dl = dprev[1][0];
#pragma unroll
                    for (int jj = 0; jj < SKB; jj++) {                /* row 2L+1 first: lower neighbours from the lane below (dn) */
                        const int j = SKB - 1 - jj;
                        double v = op.a[1][j];
                        v = v - op.cx[1][j] * dl;                     /* (aPlusX[i]*precon[i]) * dst[i+1]       F3:514 */
                        v = v - op.cy[1][j] * dn[j];                  /* (aPlusY[i]*precon[i]) * dst[i+w]       F3:517 */
                        d1[j] = v * op.pr[1][j];
                        dl = d1[j];
                    }
                    dl = dprev[0][0];
#pragma unroll
                    for (int jj = 0; jj < SKB; jj++) {                /* row 2L: lower neighbours are this macro-step's row 2L+1 */
                        const int j = SKB - 1 - jj;
                        double v = op.a[0][j];
                        v = v - op.cx[0][j] * dl;
                        v = v - op.cy[0][j] * d1[j];
                        d0[j] = v * op.pr[0][j];
                        dl = d0[j];
                        xchg_store2(to_ring, xg + j, xo + (unsigned)(j * sizeof(double)), d0[j]);
                        dnn[j] = xchg_load(xr + (unsigned)(j * sizeof(double)));
                    }
                }
                if (!RINGFEED && s + 1 < UB) {                        /* edge values held in registers (global boundary row) */
#pragma unroll
                    for (int j = 0; j < SKB; j++) dnn[j] = is_in_feed ? ecur[(s + 1) * SKB + j] : dnn[j];
                }
                if (WITH_DOT) {
                    /* inactive rows and padding columns hold +0.0 and rr is zero there: adding them is exact */
#pragma unroll
                    for (int j = 0; j < SKB; j++) { acc += d0[j] * rr0[j]; acc += d1[j] * rr1[j]; }
                }
#pragma unroll
                for (int j = 0; j < SKB; j++) { dprev[0][j] = d0[j]; dprev[1][j] = d1[j]; dn[j] = dnn[j]; }
                double* const pd = pdst + DIR * s * (32 * SKB);
                store_cells(pd, d0);
                store_cells(pd + 33 * SKB, d1);
                if (GOUT) {
                    const int xb = (t0 + DIR * s - E_OUT) * SKB;      /* first column of the edge lane's block */
#pragma unroll
                    for (int j = 0; j < SKB; j++)
                        st_global_if(is_out && (unsigned)(xb + j) < (unsigned)w, po_lane + DIR * s * SKB + j, DIR > 0 ? d1[j] : d0[j]);
                }
                if (RINGFEED) ring_rearm(in_rest + (unsigned)(s * SKB * (int)sizeof(double)), sent_hi);
                if (s == UB - 1) {                                    /* every lane is past tile ordinal b */
                    mbar_arrive_if(lane0, empty0 + 8u * rel_pos);
                    rel_pos = pos_step(rel_pos);
                }
            }
        };

        double eb0[SW_GV], eb1[SW_GV];
#pragma unroll
        for (int j = 0; j < SW_GV; j++) { eb0[j] = 0.0; eb1[j] = 0.0; }
        constexpr std::false_type F{}; constexpr std::true_type T{};
        const int nsteps = nblk * UB;
        auto credit_wait = [&](const int last_step) {
            for (unsigned spins = 0; last_step - s_credit[warp] >= SW_RB && spins < (1u << 24); spins++) {}
        };
        if (ring_feed) {
            /* prologue: the producer's macro-steps 8..15 (one slot group); only the last one carries a column of mine */
            const unsigned a = inbox0 + (unsigned)((D - 7) * SKB * (int)sizeof(double));
            double e0[SKB];
            for (int i = 0; i < 8; i++) {
                ring_peek(a + (unsigned)(i * SKB * (int)sizeof(double)), e0);
                ring_wait(a + (unsigned)(i * SKB * (int)sizeof(double)), e0, true, true);
                ring_rearm(a + (unsigned)(i * SKB * (int)sizeof(double)), sent_hi);
            }
#pragma unroll
            for (int j = 0; j < SKB; j++) dn[j] = is_in ? e0[j] : 0.0;
            ring_peek(inbox0 + (unsigned)((D + 1) * SKB * (int)sizeof(double)), epk[1]);            /* macro-step 1 */
        }
        const int last_ring_block = nblk - 2;             /* the producer's last published macro-step is taken in this block */
        auto do_block = [&](const int b, double (&ecur)[SW_GV], double (&enext)[SW_GV]) {
            const int t0 = tfirst + DIR * b * UB;
            if (trace && b == 1) tr_first = (long long)globaltimer_ns();
            const int sc0 = b * UB;
            const int cin = t0 - E_IN;
            if (out_ring && (b & 1) == 0) credit_wait(sc0 + 2 * UB - 1);
            if (use_feed && !ring_feed) {                              /* global boundary row: block by block, one block ahead */
                feed.fetch(enext, cin + DIR * UB, sc0 + UB, DIR);
                bool ok = true;
#pragma unroll
                for (int i = 0; i < SW_GV; i++) ok = ok && !is_sentinel_hi(ecur[i]);
                if (!__all_sync(HALF, ok)) {
                    for (unsigned spins = 0;; spins++) {
                        ok = true;
#pragma unroll
                        for (int i = 0; i < SW_GV; i++) ok = ok && !is_sentinel_hi(ecur[i]);
                        ok = ok || spins > (1u << 24);
                        if (__all_sync(HALF, ok)) break;
#pragma unroll
                        for (int s = 0; s < SW_FG; s++)
#pragma unroll
                            for (int j = 0; j < SKB; j++)
                                if (is_sentinel_hi(ecur[s * SKB + j]))
                                    ecur[s * SKB + j] = ld_volatile(feed.src + feed.index((cin + DIR * s) * SKB + j, sc0 + s, j));
                    }
#pragma unroll
                    for (int i = 0; i < SW_GV; i++) if (is_sentinel_hi(ecur[i])) ecur[i] = 0.0;     /* gave up: dead producer */
                }
            }
            const bool rf = ring_feed && b <= last_ring_block;
            if (out_ring) {
                /* the edge lane's results go straight into the next strip's ring (from block 1 on) */
                const bool to_ring = is_out && b >= 1;
                double* po_lane = to_ring ? out_base + (sc0 & (SW_RB - 1)) * SKB : oscratch;
                if (rf) run_block(T, F, t0, sc0, ecur, po_lane, to_ring);
                else    run_block(F, F, t0, sc0, ecur, po_lane, to_ring);
            } else {
                double* po_lane = out_base + (long long)(t0 - E_OUT) * SKB;   /* only dereferenced under its predicate */
                if (rf) run_block(T, T, t0, sc0, ecur, po_lane, false);
                else    run_block(F, T, t0, sc0, ecur, po_lane, false);
            }
            pdst += DIR * UB * (32 * SKB);
            st_int_if(pub_credit, credit_of_prev, (b + 1) * UB + D);   /* I have consumed my producer's steps up to here */
        };
        if (use_feed && !ring_feed) feed.fetch(eb0, tfirst - E_IN, 0, DIR);
        for (int b = 0; b < nblk; b += 2) {
            do_block(b, eb0, eb1);
            if (b + 1 < nblk) do_block(b + 1, eb1, eb0);
        }
        if (out_ring) {
            /* 8 more zero pairs: the consumer takes one slot per macro-step through its block nblk-2 */
            credit_wait(nsteps + UB - 1);
            if (is_out) {
                double* p = out_base + (nsteps & (SW_RB - 1)) * SKB;
                for (int i = 0; i < UB * SKB; i++) p[i] = 0.0;
            }
        }
        if (trace && lane == 0) {
            long long* tr = trace + (long long)k * 8;
            tr[0] = tr_enter; tr[1] = tr_first; tr[2] = 0; tr[3] = (long long)globaltimer_ns();
            tr[4] = 0; tr[5] = nblk; tr[6] = slot; tr[7] = (long long)get_smid();
        }
    }
    if (WITH_DOT && live && !loader) {
        double ws = warp_sum(acc);
        if (lane == 0) {
            strip_part[k] = ws;
            __threadfence();
            unsigned prev = atomicAdd(&sc->ctr_b, 1u);
            if (prev == (unsigned)nown - 1u) {
                __threadfence();
                if (nown == g.nstrips) {
                    double total = 0.0;
                    for (int kk = 0; kk < g.nstrips; kk++) total += ld_cg(strip_part + kk);
                    if (dot_mode != DOT_NONE) apply_dot_result(sc, dot_mode, total);
                }
                sc->ctr_b = 0;
            }
        }
    }
    cluster.sync();
}
