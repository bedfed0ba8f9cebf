// solids.cuh -- one SOLIDS interval (accumulation and washoff of solids on impervious land) for one segment.
//
// Behaviour: HSPF HIMPSLD as HSP2 runs it, src/hsp2/hsp2/SOLIDS.py:88-140 (_solids_ loop body).  Metric runs: suro / surs arrive
// as the reference's `SURO * 0.0394` of the stored series, ACCSDP is converted where it is used (:84-87), the stored series by save().
// Same transport-capacity washoff as SEDMNT's (method 1 on SURS+SURO when SDOPFG == 1, else on SURO), then the
// once-a-day accumulation / removal that only happens after a dry day (DRYDFG latch).  The lateral solids inflow
// SLSLD is read by the reference but never used (SOLIDS.py:93), so it is not an input here.
#pragma once
#include "land_common.cuh"

struct SolidsState {
    double slds;
    int drydfg;
    template <class SEG>
    __device__ __forceinline__ void load(const SEG &s) {
        slds = s.st(S_SLDS);
        drydfg = s.st(S_SLD_DRYDFG) != 0.0;
    }
    template <class SEG>
    __device__ __forceinline__ void store(const SEG &s) const {
        s.st(S_SLDS, slds);
        s.st(S_SLD_DRYDFG, drydfg ? 1.0 : 0.0);
    }
};

template <class SEG>
__device__ __forceinline__ void solids_step(const SEG &s, SolidsState &w, double suro, double surs, double prec) {
    const double delt60 = s.L.delt60;
    const bool dayfg = s.calfl & HSP2G_CAL_DAYFG;
    double sosld = 0.0;
    if (suro > 0.0) {  // SOLIDS.py:103-126
        const double keim = s.par(P_KEIM), jeim = s.par(P_JEIM);
        if (s.flag(FB_SLD_SDOP1)) {
            const double arg = surs + suro;
            const double stcap = delt60 * keim * hsp_pow(HDIV(arg, delt60), jeim);
            sosld = stcap > w.slds ? HDIV(w.slds * suro, arg) : HDIV(stcap * suro, arg);
            w.slds = w.slds - sosld;
        } else {
            const double stcap = delt60 * keim * hsp_pow(HDIV(suro, delt60), jeim);
            if (stcap > w.slds) {
                sosld = w.slds;
                w.slds = 0.0;
            } else {
                sosld = stcap;
                w.slds = w.slds - sosld;
            }
        }
    }
    if (!dayfg) {  // SOLIDS.py:129-142
        if (prec > 0.0) w.drydfg = 0;
    } else {
        if (w.drydfg) {
            double accsdp = s.daily(FB_M_ACCSDP, M_ACCSDP, P_ACCSDP);
            if (s.metric()) accsdp = HDIV(accsdp * 1.10231, 2.471);  // SOLIDS.py:85: tonnes/ha to tons/ac
            w.slds = accsdp + w.slds * (1.0 - s.daily(FB_M_REMSDP, M_REMSDP, P_REMSDP));
        }
        w.drydfg = prec > 0.0 ? 0 : 1;
    }
    s.save(O_SOLIDS_SOSLD, sosld);
    s.save(O_SOLIDS_SLDS, w.slds);
}
