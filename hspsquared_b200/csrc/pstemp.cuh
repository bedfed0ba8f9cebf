// pstemp.cuh -- one PSTEMP interval (soil-layer temperatures) for one segment, English units.
//
// Behaviour: HSPF HPERTMP as HSP2 runs it, src/hsp2/hsp2/PSTEMP.py:130-215 (_pstemp_ loop body) with the array
// pre-scalings of :120-126 applied per interval.  From the behavioural spec: the literal 0.555
// (and 0.5555 for the TSOPFG=1 tables), `airts` frozen at AIRTMP[0] when ATEMP is active, TSOPFG compared as a
// number, clamps at +-100 C with error counts.
#pragma once
#include "land_common.cuh"

struct PstempState {
    double sltmp, ultmp, lgtmp, airts;
    template <class SEG>
    __device__ __forceinline__ void load(const SEG &s) {
        sltmp = s.st(S_SLTMP); ultmp = s.st(S_ULTMP); lgtmp = s.st(S_LGTMP); airts = s.st(S_AIRTS);
    }
    template <class SEG>
    __device__ __forceinline__ void store(const SEG &s) const {
        s.st(S_SLTMP, sltmp); s.st(S_ULTMP, ultmp); s.st(S_LGTMP, lgtmp); s.st(S_AIRTS, airts);
    }
};

struct SoilTemps {  // deg F, what PWTGAS reads from ts
    double sltmp, ultmp, lgtmp;
};

template <bool HOURLY, class SEG>
__device__ __forceinline__ void pstemp_step(const SEG &s, PstempState &w, double airtmp, SoilTemps &o) {
    const bool hrfg = HOURLY ? true : (s.calfl & HSP2G_CAL_HRFG);
    const double airtcs = (w.airts - 32.0) * 0.555;
    const double airtc = (airtmp - 32.0) * 0.555;
    if (hrfg) {
        const double aslt = (s.daily(FB_M_ASLT, M_ASLT, P_ASLT) - 32.0) * 0.555;
        const double bslt = s.daily(FB_M_BSLT, M_BSLT, P_BSLT);
        w.sltmp = aslt + bslt * airtc;
    }
    if (s.flag(FB_TSOP1)) {
        if (hrfg) {
            const double ault = (s.daily(FB_M_ULTP1, M_ULTP1, P_ULTP1) - 32.0) * 0.5555;
            const double bult = s.daily(FB_M_ULTP2, M_ULTP2, P_ULTP2);
            w.ultmp = ault + bult * airtc;
            w.lgtmp = (s.daily(FB_M_LGTP1, M_LGTP1, P_LGTP1) - 32.0) * 0.5555;
        }
    } else {
        const double ulsmo = s.daily(FB_M_ULTP1, M_ULTP1, P_ULTP1);
        const double ultdif = 0.555 * s.daily(FB_M_ULTP2, M_ULTP2, P_ULTP2);
        const double ultmps = w.ultmp;
        w.ultmp = ultmps + ulsmo * (airtcs + ultdif - ultmps);
        const double lgsmo = s.daily(FB_M_LGTP1, M_LGTP1, P_LGTP1);
        const double lgtdif = 0.555 * s.daily(FB_M_LGTP2, M_LGTP2, P_LGTP2);
        const double lgtmps = w.lgtmp;
        if (s.flag(FB_TSOP0))
            w.lgtmp = lgtmps + lgsmo * (airtcs + lgtdif - lgtmps);
        else
            w.lgtmp = lgtmps + lgsmo * (w.ultmp + lgtdif - lgtmps);
    }
    int e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0, e5 = 0;
    if (w.sltmp < -100.0) { w.sltmp = -100.0; e0 = 1; }
    if (w.sltmp > 100.0) { w.sltmp = 100.0; e3 = 1; }
    if (w.ultmp < -100.0) { w.ultmp = -100.0; e1 = 1; }
    if (w.ultmp > 100.0) { w.ultmp = 100.0; e4 = 1; }
    if (w.lgtmp < -100.0) { w.lgtmp = -100.0; e2 = 1; }
    if (w.lgtmp > 100.0) { w.lgtmp = 100.0; e5 = 1; }
    if (!s.flag(FB_AIRTFG)) w.airts = airtmp;

    o.sltmp = (w.sltmp * 9.0 / 5.0) + 32.0;
    o.ultmp = (w.ultmp * 9.0 / 5.0) + 32.0;
    o.lgtmp = (w.lgtmp * 9.0 / 5.0) + 32.0;
    if (e0 | e1 | e2 | e3 | e4 | e5) {
        s.count(E_PSTEMP0, e0);
        s.count(E_PSTEMP1, e1);
        s.count(E_PSTEMP2, e2);
        s.count(E_PSTEMP3, e3);
        s.count(E_PSTEMP4, e4);
        s.count(E_PSTEMP5, e5);
    }
    s.save(O_PSTEMP_AIRTC, (airtc * 9.0 / 5.0) + 32.0);
    s.save(O_PSTEMP_SLTMP, o.sltmp);
    s.save(O_PSTEMP_ULTMP, o.ultmp);
    s.save(O_PSTEMP_LGTMP, o.lgtmp);
}
