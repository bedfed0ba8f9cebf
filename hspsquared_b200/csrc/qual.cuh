// qual.cuh -- one interval of a quality constituent of a pervious (PQUAL) or impervious (IQUAL) land segment.
//
// Behaviour: HSPF HPERQUA / HIMPQUA as HSP2 runs them, src/hsp2/hsp2/PQUAL.py:208-395 and IQUAL.py:189-282 (loop bodies of
// _pqual_ / _iqual_), English units.  The reference loops constituents inside one segment; constituents do not interact, so
// here one ROW of a batch is one (segment, constituent) pair: its inputs are the segment's flow / sediment / solids outflow
// series, its parameters the constituent's QUAL-PROPS / QUAL-INPUT tables (host side: segstore.QualStore).
//   - removal with detached and scoured sediment (qualsd) / with solids washoff (washsd): potency factors;
//   - accumulation on the surface and washoff by overland flow (qualof / washof): QSOFG 1 (daily accumulation), 2 (per
//     interval, with atmospheric deposition and the removal rate raised to delt60/24) and IQUAL's -1 (constant concentration);
//   - fixed concentrations in interflow and groundwater outflow (qualif / qualgw), PQUAL only.
// The lateral inflow name SLIQO is rebound to an all-zero output array in both kernels (PQUAL.py:205-206,222; IQUAL.py:174,
// 195), so it is the constant 0 here and the SLIQO / INFLOW outputs are zeros.
#pragma once
#include "land_common.cuh"

struct QualState {
    double sqo, soqsp;  // carried: surface storage; IQUAL's last computed washoff potency
    // the constituent's parameters for the current day (monthly tables are interpolated to the day, constant within it)
    double wsfac, potfw, potfs, acqop, remqop, ioqc, aoqc, adfx, adcn;
    double slifac, ilifac, alifac;  // lateral-inflow weights: constant, read once per task instead of once per interval
    template <class SEG>
    __device__ __forceinline__ void load(const SEG &s) {
        sqo = s.st(S_Q_SQO);
        soqsp = s.st(S_Q_SOQSP);
        wsfac = s.par(P_Q_WSFAC);
        slifac = s.par(P_SLIFAC);
        ilifac = s.par(P_ILIFAC);
        alifac = s.par(P_ALIFAC);
    }
    template <class SEG>
    __device__ __forceinline__ void store(const SEG &s) const {
        s.st(S_Q_SQO, sqo);
        s.st(S_Q_SOQSP, soqsp);
    }
    template <class SEG>
    __device__ __forceinline__ void refresh_daily(const SEG &s) {
        potfw = s.daily(FB_M_Q_POTFW, M_Q_POTFW, P_Q_POTFW);
        potfs = s.daily(FB_M_Q_POTFS, M_Q_POTFS, P_Q_POTFS);
        acqop = s.daily(FB_M_Q_ACQOP, M_Q_ACQOP, P_Q_ACQOP);
        remqop = s.daily(FB_M_Q_REMQOP, M_Q_REMQOP, P_Q_REMQOP);
        ioqc = s.daily(FB_M_Q_IOQC, M_Q_IOQC, P_Q_IOQC);
        aoqc = s.daily(FB_M_Q_AOQC, M_Q_AOQC, P_Q_AOQC);
        adfx = s.daily(FB_M_Q_ADFX, M_Q_ADFX, P_Q_ADFX);
        adcn = s.daily(FB_M_Q_ADCN, M_Q_ADCN, P_Q_ADCN);
    }
};

template <class SEG>
__device__ __forceinline__ void pqual_step(const SEG &s, QualState &w) {
    const double delt60 = s.L.delt60;
    const bool dayfg = s.calfl & HSP2G_CAL_DAYFG;
    const double suro = s.forcing(F_SURO), ifwo = s.forcing(F_IFWO), agwo = s.forcing(F_AGWO), pero = s.forcing(F_PERO);
    double suroqs = 0.0, soqsp = -1.0e30, scrqs = 0.0, washqs = 0.0, soqs = 0.0;
    if (s.flag(FB_Q_QSDFG)) {  // qualsd, PQUAL.py:250-276
        const double wssd = s.forcing(F_WSSD), scrsd = s.forcing(F_SCRSD), sliqsp = s.forcing(F_SLIQSP);
        const double slifac = w.slifac;
        if (wssd == 0.0)
            washqs = 0.0;
        else if (sliqsp >= 0.0)
            washqs = wssd * (sliqsp * slifac + w.potfw * (1.0 - slifac));
        else
            washqs = wssd * w.potfw;
        scrqs = scrsd == 0.0 ? 0.0 : scrsd * w.potfs;
        soqs = washqs + scrqs;
        const double lsosed = wssd + scrsd;
        soqsp = lsosed > 0.0 ? HDIV(soqs, lsosed) : -1.0e30;
        suroqs = soqs;
    }
    double suroqo = 0.0, adtot = 0.0, adfxfx = 0.0, adcnfx = 0.0, soqoc = -1.0e30, soqo = 0.0;
    if (s.flag(FB_Q_QSO_NZ)) {  // qualof, PQUAL.py:284-325
        if (dayfg && s.flag(FB_Q_QSO_EQ1)) w.sqo = w.acqop + w.sqo * (1.0 - w.remqop);
        adfxfx = s.has_forcing(F_QADFX) ? s.forcing(F_QADFX) : w.adfx;
        adcnfx = (s.has_forcing(F_QADCN) ? s.forcing(F_QADCN) : w.adcn) * s.forcing(F_PREC) * 3630.0;
        adtot = adfxfx + adcnfx;
        const double intot = adtot + 0.0;
        if (s.flag(FB_Q_QSO_EQ2)) {
            double dummy = w.remqop + HDIV(intot, HDIV(w.acqop, w.remqop));
            if (dummy > 1.0) dummy = 1.0;
            w.sqo = w.acqop * HDIV(delt60, 24.0) + w.sqo * hsp_pow(1.0 - dummy, HDIV(delt60, 24.0));
        }
        w.sqo = w.sqo + intot;
        if (suro > 0.0 && w.sqo > 0.0) {
            double dummy = suro * w.wsfac;
            if (!(dummy < 1.0e-5)) {
                dummy = 1.0 - hsp_exp(-dummy);
                soqo = w.sqo * dummy;
                w.sqo = w.sqo - soqo;
            }
        }
        soqoc = suro > 0.0 ? HDIV(soqo, suro) : -1.0e30;
        suroqo = soqo;
    }
    const double soqual = suroqs + suroqo;
    double poqual = soqual, ioqual = 0.0, aoqual = 0.0;
    const double soqc = suro > 0.0 ? HDIV(soqual, suro) : -1.0e30;
    if (s.flag(FB_Q_QIFWFG)) {  // qualif, PQUAL.py:339-353
        double ioqc = w.ioqc * 3630.0;
        if (s.flag(FB_Q_VIQC34)) ioqc = ioqc * 6.238e-5;
        if (ifwo > 0.0) {
            const double iliqc = s.forcing(F_ILIQC), ilifac = w.ilifac;
            const double ioqce = iliqc >= 0.0 ? iliqc * ilifac + ioqc * (1.0 - ilifac) : ioqc;
            ioqual = ioqce * ifwo;
        }
        poqual = poqual + ioqual;
    }
    if (s.flag(FB_Q_QAGWFG)) {  // qualgw, PQUAL.py:356-371
        double aoqc = w.aoqc * 3630.0;
        if (s.flag(FB_Q_VAQC34)) aoqc = aoqc * 6.238e-5;
        if (agwo > 0.0) {
            const double aliqc = s.forcing(F_ALIQC), alifac = w.alifac;
            const double aoqce = aliqc >= 0.0 ? aliqc * alifac + aoqc * (1.0 - alifac) : aoqc;
            aoqual = aoqce * agwo;
        }
        poqual = poqual + aoqual;
    }
    const double poqc = pero > 0.0 ? HDIV(poqual, pero) : -1.0e30;
    s.save(O_PQUAL_SOQUAL, soqual);
    s.save(O_PQUAL_IOQUAL, ioqual);
    s.save(O_PQUAL_AOQUAL, aoqual);
    s.save(O_PQUAL_POQUAL, poqual);
    s.save(O_PQUAL_SQO, w.sqo);
    s.save(O_PQUAL_SOQSP, soqsp);
    s.save(O_PQUAL_SOQOC, soqoc > -1.0 ? HDIV(soqoc, 3630.0) : soqoc);  // PQUAL.py:379-382
    s.save(O_PQUAL_SOQC, HDIV(soqc, 3630.0));                          // the sentinel is divided too (:383)
    s.save(O_PQUAL_IOQC, ifwo > 0.0 ? HDIV(HDIV(ioqual, ifwo), 3630.0) : -1.0e30);
    s.save(O_PQUAL_AOQC, agwo > 0.0 ? HDIV(HDIV(aoqual, agwo), 3630.0) : -1.0e30);
    s.save(O_PQUAL_POQC, pero > 0.0 ? HDIV(poqc, 3630.0) : -1.0e30);
    s.save(O_PQUAL_WASHQS, washqs);
    s.save(O_PQUAL_SCRQS, scrqs);
    s.save(O_PQUAL_SOQS, soqs);
    s.save(O_PQUAL_SOQO, soqo);
    s.save(O_PQUAL_PQADWT, adcnfx);
    s.save(O_PQUAL_PQADDR, adfxfx);
    s.save(O_PQUAL_PQADEP, adtot);
    s.save(O_PQUAL_SLIQO, 0.0);
    s.save(O_PQUAL_INFLOW, 0.0);
}

template <class SEG>
__device__ __forceinline__ void iqual_step(const SEG &s, QualState &w) {
    const double delt60 = s.L.delt60;
    const bool dayfg = s.calfl & HSP2G_CAL_DAYFG;
    const double suro = s.forcing(F_SURO);
    double suroqs = 0.0, soqs = 0.0;
    if (s.flag(FB_Q_QSDFG)) {  // washsd, IQUAL.py:201-217
        const double sosld = s.forcing(F_SOSLD);
        if (sosld != 0.0) {
            const double sliqsp = s.forcing(F_SLIQSP);
            if (sliqsp >= 0.0) {
                const double slifac = w.slifac;
                w.soqsp = sliqsp * slifac + w.potfw * (1.0 - slifac);
                soqs = sosld * w.soqsp;
            } else {
                w.soqsp = w.potfw;
                soqs = sosld * w.potfw;
            }
        }
        suroqs = soqs;
    }
    double suroqo = 0.0, adtot = 0.0, adfxfx = 0.0, adcnfx = 0.0, soqo = 0.0, soqoc = 0.0;
    if (s.flag(FB_Q_QSO_GE1)) {  // washof, IQUAL.py:226-257
        if (dayfg && s.flag(FB_Q_QSO_EQ1)) w.sqo = w.acqop + w.sqo * (1.0 - w.remqop);
        adfxfx = s.has_forcing(F_QADFX) ? s.forcing(F_QADFX) : w.adfx;
        adcnfx = (s.has_forcing(F_QADCN) ? s.forcing(F_QADCN) : w.adcn) * s.forcing(F_PREC) * 3630.0;
        adtot = adfxfx + adcnfx;
        if (s.flag(FB_Q_QSO_EQ2)) {
            double dummy = w.remqop + HDIV(adtot + 0.0, HDIV(w.acqop, w.remqop));
            if (dummy > 1.0) dummy = 1.0;
            w.sqo = w.acqop * HDIV(delt60, 24.0) + w.sqo * hsp_pow(1.0 - dummy, HDIV(delt60, 24.0));
        }
        w.sqo = w.sqo + 0.0 + adtot;
        if (suro > 0.0 && w.sqo > 0.0) {
            soqo = w.sqo * (1.0 - hsp_exp(-suro * w.wsfac));
            w.sqo = w.sqo - soqo;
        }
        soqoc = suro > 0.0 ? HDIV(HDIV(soqo, suro), 3630.0) : -1.0e30;
        suroqo = soqo;
    } else if (s.flag(FB_Q_QSO_M1)) {  // constant concentration, IQUAL.py:259-269
        const double acqop = w.acqop * 0.2266;
        soqo = suro * acqop;
        soqoc = HDIV(acqop, 3630.0);
        w.sqo = 0.0;
        suroqo = soqo;
    }
    const double soqual = suroqs + suroqo;
    s.save(O_IQUAL_SOQUAL, soqual);
    s.save(O_IQUAL_SOQC, suro > 0.0 ? HDIV(HDIV(soqual, suro), 3630.0) : -1.0e30);
    s.save(O_IQUAL_SQO, w.sqo);
    s.save(O_IQUAL_SOQS, soqs);
    s.save(O_IQUAL_SOQOC, soqoc);
    s.save(O_IQUAL_SOQO, soqo);
    s.save(O_IQUAL_SOQSP, w.soqsp);
    s.save(O_IQUAL_IQADWT, adcnfx);
    s.save(O_IQUAL_IQADDR, adfxfx);
    s.save(O_IQUAL_IQADEP, adtot);
    s.save(O_IQUAL_SLIQO, 0.0);
    s.save(O_IQUAL_INFLOW, 0.0);
}

// The time loop of a warp of (segment, constituent) rows; the __global__ wrappers (qual_kernel.cu, hspsquared_b200/jit.py)
// only fix the template arguments.
template <bool IMP, class IO, class LAY>
__device__ __forceinline__ void qual_body(const LandLaunch &L) {
    LAND_SMEM_DECL;
    int lane = threadIdx.x & (LAND_TILE - 1);
    const int wib = threadIdx.x / LAND_TILE;
    HSP_PIN32(lane);  // kept in a register: rematerialising it costs an S2R per use
    const int nf = IO::STATIC ? IO::NF : L.nf;
    typedef Seg<IO, 0ull, false, DynFlags, LAY> SegT;
    PipeT<false, LAY> pipe;
    pipe.init(L, smem, wib, lane, nf, 0, 0);
    Sched sched(L, lane, wib);
    int tile, t0, t1;
    while (sched.next(tile, t0, t1)) {
        pipe.begin_task(tile, t0, t1);
        SegT s(L, tile, lane, t0, pipe.hot_area());
        QualState q;
        q.load(s);
        for (int t = t0; t < t1; ++t) {
            s.begin_step(pipe.acquire());
            if (t == t0 || (s.calfl & HSP2G_CAL_NEWDAY)) q.refresh_daily(s);
            if (IMP)
                iqual_step(s, q);
            else
                pqual_step(s, q);
            s.end_step();
            pipe.release(t);
        }
        q.store(s);
        pipe.end_task();
        sched.done(tile);
    }
}
