// snde_team.cuh -- the strip program of snde_strip.cuh run by a TEAM OF FIVE WARPS per strip instead of a pair:
//   G1  geometry forward   row t+1: reproject + start the tap gathers;  row t: blend -> warped row + tap derivatives
//   G2  geometry backward  row t-3: dL/d(warped) x tap derivatives, back through the projection -> d/d(depth), d/dP;
//                                   the owner's smoothness pairs (first pass)
//   S0, S1, S2  SSIM of colour channel 0, 1, 2: rolling window sums, error value, backward coefficients, their box sums
//               -> dL/d(warped) of the channel; each brings its channel's owner rows into the ring (S0 also depth and y)
// Same rings, same arithmetic, same task table, ONE named barrier of 160 threads per row.  Why: in the pair kernel each warp
// carries the state of three channels or of both geometry directions and sits at the 168-register limit of 12 warps per SM;
// here the state is spread over five lighter warps, the kernel runs 20 warps per SM (five per scheduler) and the roles get
// register budgets that fit them (setmaxnreg: geometry 128, SSIM channels 72 -- snde_team.cu).
// Lean fast-path instances only (<GRAD, FAST> with or without DISP); everything else stays on the pair kernel.
#pragma once
#include "snde_strip.cuh"

namespace snde {

struct TeamCfg {
    static constexpr int TEAMS = 4;            // strips in flight per CTA: warp j of every role group serves team j (same scheduler)
    static constexpr int ROLES = 5;
    static constexpr int NT = TEAMS * ROLES * 32;
};

// ===============================================================================================================
// SSIM of one colour channel
// ===============================================================================================================
template <bool GRAD, bool DISP>
SNDE_STRIP_FN void team_ssim_role(const DevLoss &p, int task, int who, PairSmem &sm, int bar, int ch) {
    const int lane = who & 31;
    StripTask T;
    decode_task(p, task, lane, T);
    const DevScale &sc = p.sc[T.scale];
    OwnerRows<true> rows; rows.init(p, sc, T);
    const bool lead = ch == 0;                       // channel 0's warp also brings depth and meshgrid y
    float alpha = p.alpha, oma = p.oma, g1 = p.oma * sc.gphoto;
    SNDE_PIN_F32(alpha); SNDE_PIN_F32(oma); SNDE_PIN_F32(g1);
    float ngSm[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) ngSm[q] = T.colB[q] ? -(-0.5f * p.alpha * sc.gphoto) : 0.f;
    const int y0 = T.y0, y1 = T.y1, h = T.h;
    for (int k = 0; k < T.n_pass; ++k) {
        HS s1, s2; CS c1, c2;
        hs_zero(s1); hs_zero(s2); cs_zero(c1); cs_zero(c2);
        float err_acc = 0.f;
        rows.prefetch_row_channel(sm, T, lane, T.t_begin, ch, lead); rows.prefetch_row_channel(sm, T, lane, T.t_begin + 1, ch, lead);
        rows.prefetch_row_channel(sm, T, lane, T.t_begin + 2, ch, lead);
        row_copy_wait<1>();
        if (DISP && lead) { rows.disparity_to_depth(sm, T, lane, T.t_begin); rows.disparity_to_depth(sm, T, lane, T.t_begin + 1); }
        pair_barrier(who, bar);
        for (int t = T.t_begin; t <= T.t_end + 1; ++t) {
            pair_barrier(who, bar);              // row t of the warped image is in shared memory
            if (t <= T.t_end && T.do_photo) {
                const int rb = t - 1, rc = t - 2;
                const bool doB = t >= y0, doC = t >= y0 + 2;
                const bool rowB = doB && rb >= 0 && rb < h;
                const bool row_inner = rb >= y0 && rb < y1;
                const F2 xr = sm.X[t & 7][ch][lane], yr = sm.Y[t & 3][ch][lane];
                HS H;
                {
                    const float to_right[2] = {xr.y, yr.y}, to_left[2] = {xr.x, yr.x};
                    float from_left[2], from_right[2];
                    neighbor_exchange<2>(who, to_right, to_left, from_left, from_right);
                    hsum_row(xr, yr, from_left[0], from_right[0], from_left[1], from_right[1], H);
                }
                CS co; cs_zero(co);
                {
                    HS tot; hs_add(s1, H, tot);
                    const F2 xc = sm.X[rb & 7][ch][lane], yc = sm.Y[rb & 3][ch][lane];
                    SsimP ss; ssim_sums_p(tot.x, tot.y, tot.q, tot.xy, ss);
                    const P2 v = p2fma(p2s(-0.5f), ss.S, p2s(0.5f));
                    const float vq[2] = {p2lo(v), p2hi(v)};
                    const float vcq[2] = {clampf(vq[0], 0.f, 1.f), clampf(vq[1], 0.f, 1.f)};
                    const float dq[2] = {xc.x - yc.x, xc.y - yc.y};
                    float ngq[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float err = fmaf(alpha, vcq[q], oma * fabsf(dq[q]));
                        if (rowB && row_inner) err_acc = fmaf(err, T.mI[q], err_acc);
                        ngq[q] = (rowB && vq[q] == vcq[q]) ? ngSm[q] : 0.f;
                    }
                    if (GRAD) ssim_sums_bwd_p(ss, tot.x, tot.y, p2(ngq[0], ngq[1]), co.a, co.b, co.c);
                }
                hs_add(s2, H, s1);
                s2 = H;
                if (GRAD) {
                    CS Hc;
                    {
                        const float al = p2lo(co.a), ah = p2hi(co.a), bl = p2lo(co.b), bh = p2hi(co.b), cl = p2lo(co.c), chh = p2hi(co.c);
                        const float to_right[3] = {ah, bh, chh}, to_left[3] = {al, bl, cl};
                        float from_left[3], from_right[3];
                        neighbor_exchange<3>(who, to_right, to_left, from_left, from_right);
                        Hc.a = p2add(p2(from_left[0], from_right[0]), p2s(al + ah));
                        Hc.b = p2add(p2(from_left[1], from_right[1]), p2s(bl + bh));
                        Hc.c = p2add(p2(from_left[2], from_right[2]), p2s(cl + chh));
                    }
                    {
                        CS bs; cs_add(c1, Hc, bs);
                        const F2 xf = sm.X[rc & 7][ch][lane], yf = sm.Y[rc & 3][ch][lane];
                        const P2 xv = p2f(xf), yv = p2f(yf);
                        const float d0 = xf.x - yf.x, d1 = xf.y - yf.y;
                        const float l0 = d0 > 0.f ? g1 : (d0 < 0.f ? -g1 : 0.f), l1 = d1 > 0.f ? g1 : (d1 < 0.f ? -g1 : 0.f);
                        const P2 po = p2sub(p2fma(p2add(yv, yv), bs.b, p2fma(xv, bs.c, bs.a)), p2(l0, l1));
                        if (doC) sm.PO[rc & 1][ch][lane] = f2p(po);
                    }
                    cs_add(c2, Hc, c1);
                    c2 = Hc;
                }
            }
            rows.prefetch_row_channel(sm, T, lane, t + 3, ch, lead);
            row_copy_wait<1>();
            if (DISP && lead) rows.disparity_to_depth(sm, T, lane, t + 2);
        }
        row_copy_wait<0>();
        pair_barrier(who, bar);                  // the geometry warps are through with the ring
        {
            float v[1] = {err_acc};
            warp_sum<1>(who, v);
            if (T.n_others > 0 && lane == 0 && v[0] != 0.f) atomic_add_double(sc.loss + (T.clsF ? 0 : 1), (double)v[0]);
        }
    }
}

// ===============================================================================================================
// geometry warps: both share the per-task / per-pass setup
// ===============================================================================================================
template <bool GRAD, bool DISP>
SNDE_STRIP_FN void team_geo_setup(GeoLane<GRAD, false, false, true, DISP> &L, const DevLoss &p, const DevScale &sc, const StripTask &T) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        L.xt[q] = T.in[q] ? ldg(sc.xtab + (T.ix0 + q)) : 0.f;
        L.mxs[q] = (T.colB[q] && T.ix0 + q + 1 < T.w) ? 1.f : 0.f; L.mys[q] = T.colB[q] ? 1.f : 0.f;
        L.mxi[q] = L.mxs[q] * T.mI[q]; L.myi[q] = L.mys[q] * T.mI[q];
    }
    L.xtp = p2(L.xt[0], L.xt[1]);
    const long long depth_n = T.clsF ? T.m : (long long)p.B + T.m;
    L.gd_base = GRAD ? sc.gdepth.p + depth_n * sc.gdepth.sn : nullptr; L.gd_sh = sc.gdepth.sh; L.gd_sw = sc.gdepth.sw;
    SNDE_PIN_PTR(L.gd_base);
    L.sx_acc = 0.f; L.sy_acc = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) L.dK[i] = 0.f;
    L.wout_base = nullptr; L.wout_sc = L.wout_sh = L.wout_sw = 0;
    L.gim_base = nullptr; L.gim_sc = L.gim_sh = L.gim_sw = 0;
}

template <bool GRAD, bool DISP>
SNDE_STRIP_FN void team_geo_pass_setup(GeoLane<GRAD, false, false, true, DISP> &L, const DevLoss &p, const DevScale &sc, const StripTask &T,
                                       PairSmem &sm, long long n) {
    const DevImage &oth = T.clsF ? sc.src : sc.tgt;
    L.oth_base = oth.p + (T.clsF ? n : (long long)T.b) * oth.sn; L.oth_sc = oth.sc; L.oth_sh = oth.sh; L.oth_sw = oth.sw;
    if (T.n_others > 0) {
        const float *Pn = sc.P + (n * 2 + (T.clsF ? 0 : 1)) * 12;
        const float *Kinv = sc.Kinv + 9 * T.b;
#pragma unroll
        for (int i = 0; i < 12; ++i) L.P[i] = ldg(Pn + i);
#pragma unroll
        for (int i = 0; i < 9; ++i) L.Ki[i] = ldg(Kinv + i);
        // (touch every entry before the row loop: see geo_role)
        float touch = L.xt[0] + L.xt[1];
#pragma unroll
        for (int i = 0; i < 12; ++i) touch += L.P[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) touch += L.Ki[i];
        if (touch == 1.2345e38f) sm.pad_[0] = 1;
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) L.P[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) L.Ki[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) L.dP[i] = 0.f;
    L.sy_prev = F2{0.f, 0.f};
}

// ---- G1: reproject + gather + blend ----
template <bool GRAD, bool DISP>
SNDE_STRIP_FN void team_geo_fwd_role(const DevLoss &p, int task, int who, PairSmem &sm, int bar) {
    const int lane = who & 31;
    StripTask T;
    decode_task(p, task, lane, T);
    const DevScale &sc = p.sc[T.scale];
    GeoLane<GRAD, false, false, true, DISP> L(p, sc, sm, T, lane);
    team_geo_setup<GRAD, DISP>(L, p, sc, T);
    const int t_begin = T.t_begin, t_end = T.t_end, h = T.h;
    for (int k = 0; k < T.n_pass; ++k) {
        const long long n = T.clsF ? (long long)k * p.B + T.b : T.m;
        team_geo_pass_setup<GRAD, DISP>(L, p, sc, T, sm, n);
        P2 wts[4], dgate[2]; float taps[24];
#pragma unroll
        for (int i = 0; i < 24; ++i) taps[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) wts[i] = p2s(0.f);
        dgate[0] = dgate[1] = p2s(0.f);
        pair_barrier(who, bar);           // owner rows t_begin, t_begin+1 are in the ring
        if (T.do_photo && t_begin < h) L.a_issue(sm.YT[t_begin & 7], sm.D[t_begin & 7][lane], wts, dgate, taps);
        for (int t = t_begin; t <= t_end + 1; ++t) {
            {
                const bool want_g = t >= T.y0 && t < T.y1;
                F2 yr[3], gr[6];
                L.a_finish(t, wts, dgate, taps, want_g, yr, gr);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) sm.Y[t & 3][ch][lane] = yr[ch];
                if (GRAD && want_g) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) sm.G[t & (SNDE_G_SLOTS - 1)][i][lane] = gr[i];
                }
            }
            pair_barrier(who, bar);
            L.a_issue(sm.YT[(t + 1) & 7], sm.D[(t + 1) & 7][lane], wts, dgate, taps, T.do_photo && t + 1 <= t_end && t + 1 < h);
        }
        pair_barrier(who, bar);
    }
}

// ---- G2: back-projection, smoothness, depth gradient, d/dP ----
template <bool GRAD, bool DISP>
SNDE_STRIP_FN void team_geo_bwd_role(const DevLoss &p, int task, int who, PairSmem &sm, int bar) {
    const int lane = who & 31;
    StripTask T;
    decode_task(p, task, lane, T);
    const DevScale &sc = p.sc[T.scale];
    GeoLane<GRAD, false, false, true, DISP> L(p, sc, sm, T, lane);
    team_geo_setup<GRAD, DISP>(L, p, sc, T);
    const int t_begin = T.t_begin, t_end = T.t_end;
    for (int k = 0; k < T.n_pass; ++k) {
        const long long n = T.clsF ? (long long)k * p.B + T.b : T.m;
        team_geo_pass_setup<GRAD, DISP>(L, p, sc, T, sm, n);
        pair_barrier(who, bar);
        for (int t = t_begin; t <= t_end + 1; ++t) {
            pair_barrier(who, bar);       // dL/d(warped) of row t-3 is ready
            L.template finish_row<true>(t - 3, k, who, T.do_smooth && k == 0);
        }
        pair_barrier(who, bar);
        {
            const float mine = warp_sum_scatter<12>(who, L.dP);
            const int slot = warp_sum_slot(lane);
            if (GRAD && T.n_others > 0 && !(lane & 1) && slot < 12 && mine != 0.f)
                atomic_add_double(sc.gP + (n * 2 + (T.clsF ? 0 : 1)) * 12 + slot, (double)mine);
        }
    }
    {
        float v[2] = {L.sx_acc, L.sy_acc};
        warp_sum<2>(who, v);
        if (lane == 0) {
            if (v[0] != 0.f) atomic_add_double(sc.loss + 2, (double)v[0]);
            if (v[1] != 0.f) atomic_add_double(sc.loss + 3, (double)v[1]);
        }
    }
}

}  // namespace snde
