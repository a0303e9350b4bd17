// Phase-split frontier pipeline: the same per-box algorithm as yr_box.h::process_box, cut into
// small uniform kernels so that every kernel's code fits the instruction cache and the serial
// per-box mathematics runs thread-per-box (32 boxes share one instruction stream):
//
//   tensor_step   warp (or CTA) per box: stage the tensors in shared memory, then one of
//                   kOpPrep      sum|M_p| and the coefficients of total degree <= 2
//                   kOpTrim      trimMs (:1131-1171), tensors compacted back to the pool
//                   kOpRestrict  build C(alpha,beta) per axis, restrict every tensor in place
//                                (:864-894), tensors compacted back, new sums / low coefficients
//   scalar_step   thread per box: isPoint, constant check (:1213), quadratic check
//                 (QuadraticCheck.py), zoom-loop bookkeeping (:1243-1252), BILS (:628-753),
//                 addTransform (:420-454), final-step re-entry (:1264), emit / queue subdivision
//   subdivide     warp (or CTA) per box: yr_box.h::subdivide_box (:1051-1129)
//
// A level alternates tensor_step and scalar_step over the shrinking list of boxes that still need
// work ("rounds"); the host reads one counter per round.  Compared with the fused kernel a box's
// tensors cross HBM once per step instead of once per level -- irrelevant while the solver is far
// from the HBM roof, and what buys instruction-cache residency (profiles/r01_*).
#pragma once
#include "yr_box.h"

namespace yr {

// ---- low-order coefficient block ------------------------------------------------------------
// layout: [0] constant, [1..D] linear, [D+1..2D] pure T_2, then cross terms (a<b) in lexicographic order
template <int D>
YR_HD int low_cross_index(int a, int b) {            // a < b
    int idx = 2 * D + 1;
    for (int i = 0; i < a; ++i) idx += D - 1 - i;
    return idx + (b - a - 1);
}

template <int D>
YR_HD void gather_low(const BoxShared<D>& sh, const double* ws, int p, double* low) {
    int idx[D];
    for (int d = 0; d < D; ++d) idx[d] = 0;
    low[0] = coef_at<D>(sh, ws, p, idx);
    for (int a = 0; a < D; ++a) {
        idx[a] = 1; low[1 + a] = coef_at<D>(sh, ws, p, idx);
        idx[a] = 2; low[1 + D + a] = coef_at<D>(sh, ws, p, idx);
        idx[a] = 0;
    }
    for (int a = 0; a < D; ++a)
        for (int b = a + 1; b < D; ++b) {
            idx[a] = 1; idx[b] = 1;
            low[low_cross_index<D>(a, b)] = coef_at<D>(sh, ws, p, idx);
            idx[a] = 0; idx[b] = 0;
        }
}

// quadratic_check on the low-order block (same dispatch as quadratic_check_box)
template <int D>
YR_HD bool quadratic_check_low(const double* l, double sumabs, double tol) {
    if (D == 2) {
        const double c[6] = {l[0], l[1], l[2], l[3], l[5], l[4]};
        return quad_check_2d(c, sumabs, tol);
    } else if (D == 3) {
        const double c[10] = {l[0], l[1], l[2], l[3], l[7], l[8], l[9], l[4], l[5], l[6]};
        return quad_check_3d(c, sumabs, tol);
    } else {
        double g[D], pure[D], cross[D][D];
        double used = fabs(l[0]);
        for (int a = 0; a < D; ++a) { g[a] = l[1 + a]; pure[a] = l[1 + D + a]; used += fabs(g[a]) + fabs(pure[a]); }
        for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) cross[a][b] = 0.0;
        for (int a = 0; a < D; ++a) for (int b = a + 1; b < D; ++b) { cross[a][b] = l[low_cross_index<D>(a, b)]; used += fabs(cross[a][b]); }
        return quad_check_nd<D>(l[0], g, pure, cross, sumabs - used, tol);
    }
}

// ---- tensor step ------------------------------------------------------------------------------
// stage a box of the level into the team's shared memory (record, layout, tensors)
template <int D, class Ctx>
YR_DEVN void stage_for_step(Ctx& cx, const PhaseArgs<D>& pa, BoxShared<D>& sh, double* ws, unsigned int b, bool subdivide) {
    const SolverParams& prm = pa.lv.prm;
    {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&pa.lv.recs_in[b]);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&sh.rec);
        for (int w = cx.tid; w < (int)(sizeof(BoxRec<D>) / 8); w += cx.nthreads) dst[w] = src[w];
    }
    cx.sync();
    if (cx.tid == 0) {
        for (int p = 0; p < D; ++p) {
            int st = 1;
            for (int a = D - 1; a >= 0; --a) { sh.shape[p][a] = sh.rec.shape[p][a]; sh.stride[p][a] = st; st *= sh.rec.shape[p][a]; }
            sh.err[p] = sh.rec.err[p];
        }
        compute_layout<D>(sh, prm.exact != 0, subdivide);
        sh.action = kActNone;
        sh.trimmed_any = 0;
    }
    cx.sync();
    cx.copy_in(ws, pa.lv.data_in + sh.rec.data_off, sh.lay.total);
    cx.sync();
}

// tensors (views in the workspace) -> packed row-major at the box's place in the pool
template <int D, class Ctx>
YR_DEVN void write_back_compact(Ctx& cx, const PhaseArgs<D>& pa, BoxShared<D>& sh, const double* ws) {
    double* dst = const_cast<double*>(pa.lv.data_in) + sh.rec.data_off;
    for (int p = 0; p < D; ++p) {
        const double* M = ws + sh.lay.base[p];
        int nrows = 1;
        for (int a = 0; a < D - 1; ++a) nrows *= sh.shape[p][a];
        const int len = sh.shape[p][D - 1];
        if (len >= 16 || D == 1) {
            for (int r = cx.warp; r < nrows; r += cx.nwarps) {
                const double* row = M + row_offset<D>(r, sh.shape[p], sh.stride[p]);
                for (int j = cx.lane; j < len; j += cx.lanes) dst[(long long)r * len + j] = row[j];
            }
        } else {
            for (int r = cx.tid; r < nrows; r += cx.nthreads) {
                const double* row = M + row_offset<D>(r, sh.shape[p], sh.stride[p]);
                for (int j = 0; j < len; ++j) dst[(long long)r * len + j] = row[j];
            }
        }
        dst += (long long)nrows * len;
    }
}

template <int D, class Ctx>
YR_DEV void tensor_step_worker(Ctx& cx, const PhaseArgs<D>& pa, BoxShared<D>& sh, double* ws) {
    const SolverParams& prm = pa.lv.prm;
    BoxRec<D>* recs = const_cast<BoxRec<D>*>(pa.lv.recs_in);
    if (cx.tid == 0) sh.c_macs = 0;
    for (;;) {
        cx.sync();
        if (cx.tid == 0) sh.work_index = cx.atomic_add(&pa.ctl->cursor, 1ull);
        cx.sync();
        const unsigned long long i = sh.work_index;
        if (i >= pa.n_list) break;
        const unsigned int b = pa.list_in ? pa.list_in[i] : (unsigned int)i;
        BoxWork<D>& wk = pa.work[b];
        const int op = pa.first_round ? (int)kOpPrep : wk.op;
        stage_for_step<D>(cx, pa, sh, ws, b, false);
        if (op == kOpPrep) {
            for (int p = 0; p < D; ++p) {
                const double s = tensor_abs_sum<D>(cx, ws + sh.lay.base[p], sh.shape[p], sh.stride[p]);
                if (cx.tid == 0) wk.sumabs[p] = s;
            }
            for (int p = cx.tid; p < D; p += cx.nthreads) gather_low<D>(sh, ws, p, wk.low[p]);
            if (cx.tid == 0) {
                for (int d = 0; d < D; ++d) { wk.cell_iv[d][0] = sh.rec.st.iv[d][0]; wk.cell_iv[d][1] = sh.rec.st.iv[d][1]; }
                wk.phase = kPhNew; wk.op = kOpNone; wk.node_level = sh.rec.level;
                wk.zoom_count = 0; wk.changed = 1; wk.should_stop = 0; wk.final_split = 0; wk.result_index = -1;
            }
        } else if (op == kOpTrim) {
            trim_box<D>(cx, sh, ws, prm);                                  // :1232
            if (sh.trimmed_any) {
                for (int p = 0; p < D; ++p) {
                    const double s = tensor_abs_sum<D>(cx, ws + sh.lay.base[p], sh.shape[p], sh.stride[p]);
                    if (cx.tid == 0) wk.sumabs[p] = s;
                }
                write_back_compact<D>(cx, pa, sh, ws);
                if (cx.tid == 0)
                    for (int p = 0; p < D; ++p) { for (int a = 0; a < D; ++a) recs[b].shape[p][a] = sh.shape[p][a]; recs[b].err[p] = sh.err[p]; }
            }
            if (cx.tid == 0) wk.op = kOpNone;
        } else if (op == kOpRestrict) {
            if (cx.tid == 0) {
                for (int d = 0; d < D; ++d) {
                    sh.alpha[d] = wk.alpha[d]; sh.beta[d] = wk.beta[d];
                    int n = 1;
                    for (int p = 0; p < D; ++p) n = sh.shape[p][d] > n ? sh.shape[p][d] : n;
                    sh.mat_axis[d] = d; sh.mat_n[d] = n;
                    sh.sumabs[d] = wk.sumabs[d];
                }
                sh.nmat = D;
            }
            cx.sync();
            restrict_system_inplace<D>(cx, sh, ws, prm);                  // :864-894
            cx.sync();
            write_back_compact<D>(cx, pa, sh, ws);
            for (int p = cx.tid; p < D; p += cx.nthreads) gather_low<D>(sh, ws, p, wk.low[p]);
            if (cx.tid == 0) {
                for (int p = 0; p < D; ++p) {
                    for (int a = 0; a < D; ++a) recs[b].shape[p][a] = sh.shape[p][a];
                    recs[b].err[p] = sh.err[p];
                    wk.sumabs[p] = sh.sumabs[p];
                }
                wk.op = kOpNone;
            }
        }
    }
    if (cx.tid == 0 && sh.c_macs) cx.atomic_add(&pa.lv.ctr->restrict_macs, sh.c_macs);
}

// ---- scalar step (thread per box) -----------------------------------------------------------------
struct ScalarCounts {
    unsigned long long boxes, zooms, dconst, dquad, dzoom, coeffs, maxlevel;
    unsigned long long t_total, t_tensor, t_extent;      // largest box queued for a tensor step by this thread
    unsigned long long s_total, s_tensor, s_extent;      // ... for subdivision
};

// The atomics differ between backends; they are passed in as a tiny policy object `At` with
//   unsigned long long add(unsigned long long* p, unsigned long long v);
template <int D, class At>
YR_HDN void scalar_step(const PhaseArgs<D>& pa, unsigned int b, ScalarCounts& cnt, At at) {
    const SolverParams& prm = pa.lv.prm;
    BoxRec<D>& rec = const_cast<BoxRec<D>*>(pa.lv.recs_in)[b];
    BoxWork<D>& wk = pa.work[b];
    IntervalState<D> st = rec.st;
    int phase = wk.phase;
    if (phase == kPhDone) return;

    auto emit = [&](uint32_t extra_flags) {
        const unsigned long long idx = at.add(&pa.lv.ctr->n_results, 1ull);
        wk.result_index = (int)(idx + pa.lv.result_base);
        if (idx < pa.lv.cap_results) {
            ResultRec<D>& r = pa.lv.results[idx];
            const bool fin = st.final_step();
            bool touch = false;
            for (int d = 0; d < D; ++d) {
                const dd A = fin ? st.pfA[d] : st.A[d], B = fin ? st.pfB[d] : st.B[d];
                dd nA; nA.hi = -A.hi; nA.lo = -A.lo;
                r.box[d][0] = dd_to_double(dd_add(B, nA));
                r.box[d][1] = dd_to_double(dd_add(B, A));
                dd nA2; nA2.hi = -st.A[d].hi; nA2.lo = -st.A[d].lo;
                const double lo = dd_to_double(dd_add(st.B[d], nA2));
                const double hi = dd_to_double(dd_add(st.B[d], st.A[d]));
                r.root[d] = fin ? (lo + hi) / 2 : (r.box[d][0] + r.box[d][1]) / 2;
                r.comb_iv[d][0] = fin ? st.pf_iv[d][0] : st.iv[d][0];
                r.comb_iv[d][1] = fin ? st.pf_iv[d][1] : st.iv[d][1];
                r.iv_lo[d] = st.iv[d][0];
                r.cell_iv[d][0] = wk.cell_iv[d][0]; r.cell_iv[d][1] = wk.cell_iv[d][1];
                r.next_mid[d] = st.next_mid[d];
                touch = touch || r.comb_iv[d][0] == wk.cell_iv[d][0] || r.comb_iv[d][1] == wk.cell_iv[d][1];
            }
            r.system = rec.system; r.parent_node = rec.parent_node; r.collector = rec.collector;
            r.level = wk.node_level;
            r.flags = extra_flags | (fin ? kResFinalStep : 0u) | (touch ? kResTouchesCell : 0u);
            r.pad = 0;
            r.path_hi = rec.path_hi; r.path_lo = rec.path_lo;
        } else at.add(&pa.lv.ctr->overflow, 1ull);
    };
    auto box_size = [&](unsigned long long& total, unsigned long long& tensor, unsigned long long& extent) {
        unsigned long long t = 0;
        for (int p = 0; p < D; ++p) {
            unsigned long long sz = 1;
            for (int a = 0; a < D; ++a) { sz *= (unsigned long long)rec.shape[p][a]; if ((unsigned long long)rec.shape[p][a] > extent) extent = rec.shape[p][a]; }
            t += sz; if (sz > tensor) tensor = sz;
        }
        if (t > total) total = t;
    };
    auto queue_tensor = [&](int op, int next_phase) {
        wk.op = op; wk.phase = next_phase;
        box_size(cnt.t_total, cnt.t_tensor, cnt.t_extent);
        const unsigned long long pos = at.add(&pa.ctl->n_next, 1ull);
        pa.list_next[pos] = b;
    };
    auto finish_box = [&]() { wk.phase = kPhDone; wk.op = kOpNone; };

    for (;;) {
        if (phase == kPhNew) {
            // ---- entry of one solvePolyRecursive invocation :1202-1224
            cnt.boxes += 1;
            if (st.is_point()) { emit(kResPointAtEntry); finish_box(); break; }
            wk.node_level += 1;
            if ((unsigned long long)wk.node_level > cnt.maxlevel) cnt.maxlevel = wk.node_level;
            for (int p = 0; p < D; ++p) { unsigned long long sz = 1; for (int a = 0; a < D; ++a) sz *= (unsigned long long)rec.shape[p][a]; cnt.coeffs += sz; }
            bool discard = false;
            if (prm.constant_check)
                for (int p = 0; p < D && !discard; ++p) {
                    const double c0 = wk.low[p][0];
                    if (fabs(c0) > wk.sumabs[p] - fabs(c0) + rec.err[p]) { discard = true; cnt.dconst += 1; }
                }
            if (!discard && ((prm.low_dim_quad_check && D <= 3) || prm.all_dim_quad_check))
                for (int p = 0; p < D && !discard; ++p)
                    if (quadratic_check_low<D>(wk.low[p], wk.sumabs[p], rec.err[p])) { discard = true; cnt.dquad += 1; }
            if (discard) { finish_box(); break; }
            if (prm.trim) { rec.st = st; queue_tensor(kOpTrim, kPhTrimmed); return; }
            phase = kPhTrimmed;
        }
        if (phase == kPhTrimmed) {
            for (int d = 0; d < D; ++d) {
                wk.entry_iv[d][0] = st.final_step() ? st.pf_iv[d][0] : st.iv[d][0];
                wk.entry_iv[d][1] = st.final_step() ? st.pf_iv[d][1] : st.iv[d][1];
                wk.last_size[d] = st.iv[d][1] - st.iv[d][0];
            }
            wk.zoom_count = 0; wk.changed = 1; wk.should_stop = 0;
        } else if (phase == kPhZoomed) {
            // after a restriction :952-954, :1249-1252
            if (st.final_step() && st.is_point()) { wk.should_stop = 1; wk.changed = 0; }
            bool all_big = true;
            for (int d = 0; d < D; ++d) {
                const double now = st.iv[d][1] - st.iv[d][0];
                all_big = all_big && (now >= wk.last_size[d] / 2);
                wk.last_size[d] = now;
            }
            if (all_big) wk.zoom_count += 1;
        }
        // ---- zoom loop :1243-1252
        bool queued = false, dead = false;
        while (wk.changed && wk.zoom_count <= prm.max_zoom_count) {
            cnt.zooms += 1;
            double lin[D][D], c0[D];
            for (int p = 0; p < D; ++p) {
                c0[p] = wk.low[p][0];
                for (int d = 0; d < D; ++d) lin[p][d] = (rec.shape[p][d] == 1) ? 0.0 : wk.low[p][1 + d];
            }
            double sub[D][2];
            BilsOut bo = bounding_interval<D>(lin, c0, wk.sumabs, rec.err, st.final_step(), sub);
            for (int d = 0; d < D; ++d)                                       // frozen axes :932-935
                if (st.iv[d][0] == st.iv[d][1]) { sub[d][0] = -1.0; sub[d][1] = 1.0; }
            bool changed = bo.changed, should_stop = bo.should_stop, thr = bo.throw_out;
            if (thr && !st.can_throw_out()) { thr = false; should_stop = true; changed = true; }   // :937-940
            wk.changed = changed ? 1 : 0; wk.should_stop = should_stop ? 1 : 0;
            if (thr) { cnt.dzoom += 1; dead = true; break; }
            if (changed) {
                double al[D], be[D];
                if (!add_transform<D>(st, sub, al, be)) { cnt.dzoom += 1; dead = true; break; }
                for (int d = 0; d < D; ++d) { wk.alpha[d] = al[d]; wk.beta[d] = be[d]; }
                rec.st = st;
                queue_tensor(kOpRestrict, kPhZoomed);
                queued = true;
                break;
            }
            // not changed: sizes are what they were, so the pass counts towards maxZoomCount (:1250)
            wk.zoom_count += 1;
        }
        if (queued) return;
        if (dead) { finish_box(); break; }
        // ---- what next? :1254-1317
        if (wk.should_stop) {
            if (st.final_step()) { emit(0u); finish_box(); break; }
            st.start_final_step();                                          // :1264 re-enter on the same tensors
            phase = kPhNew;
            continue;
        }
        if (st.final_step()) {
            st.flags |= kFlagCanThrowFinal;
            emit(kResCollector);                                            // the parent reports; its subtree supplies the points
            wk.final_split = 1;
        } else wk.final_split = 0;
        {
            const unsigned long long pos = at.add(&pa.ctl->n_sub, 1ull);
            pa.list_sub[pos] = b;
            box_size(cnt.s_total, cnt.s_tensor, cnt.s_extent);
        }
        finish_box();
        break;
    }
    rec.st = st;
}

// ---- subdivision (warp / CTA per box) -----------------------------------------------------------
template <int D, class Ctx>
YR_DEV void subdivide_step_worker(Ctx& cx, const PhaseArgs<D>& pa, BoxShared<D>& sh, double* ws) {
    if (cx.tid == 0) { sh.c_macs = 0; sh.c_subdiv = 0; }
    for (;;) {
        cx.sync();
        if (cx.tid == 0) sh.work_index = cx.atomic_add(&pa.ctl->cursor, 1ull);
        cx.sync();
        const unsigned long long i = sh.work_index;
        if (i >= pa.n_list) break;
        const unsigned int b = pa.list_in[i];
        const BoxWork<D>& wk = pa.work[b];
        stage_for_step<D>(cx, pa, sh, ws, b, true);
        if (cx.tid == 0) {
            for (int d = 0; d < D; ++d) {
                sh.sumabs[d] = wk.sumabs[d];
                sh.entry_iv[d][0] = wk.entry_iv[d][0]; sh.entry_iv[d][1] = wk.entry_iv[d][1];
            }
            sh.level = wk.node_level;
            sh.result_index = wk.result_index;
        }
        cx.sync();
        subdivide_box<D>(cx, pa.lv, sh, ws, wk.final_split != 0);
    }
    if (cx.tid == 0) {
        cx.atomic_add(&pa.lv.ctr->restrict_macs, sh.c_macs);
        cx.atomic_add(&pa.lv.ctr->subdivisions, sh.c_subdiv);
    }
}

}  // namespace yr
