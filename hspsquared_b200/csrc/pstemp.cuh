// pstemp.cuh -- one PSTEMP interval (soil-layer temperatures) for one segment.
//
// Behaviour: HSPF HPERTMP as HSP2 runs it, src/hsp2/hsp2/PSTEMP.py:130-215 (_pstemp_ loop body) with the array
// pre-scalings of :120-126 applied per interval.  From the behavioural spec: the literal 0.555
// (and 0.5555 for the TSOPFG=1 tables), `airts` frozen at AIRTMP[0] when ATEMP is active, TSOPFG compared as a
// number, clamps at +-100 C with error counts.  Metric runs (uunits == 2): the tables and states are deg C already -- no
// conversion of ASLT / ULTP1 / LGTP1 / ULTP2 / LGTP2 (PSTEMP.py:120-126,146-149) and none of the stored series (:205-215); AIRTMP
// is still taken from deg F to deg C, as an array on entry there (:110-111), i.e. the same operations on the same values.
#pragma once
#include "land_common.cuh"

struct PstempState {
    double sltmp, ultmp, lgtmp, airts;
    // the six (optionally monthly) coefficients, interpolated once per simulated day instead of on every interval
    // (PSTEMP.py:120-126 builds them as per-interval arrays that are constant within a day): shared memory
    ColdD aslt, bslt, ultp1, ultp2, lgtp1, lgtp2;

    __device__ __forceinline__ void bind(double *c) {
        ColdD *m[NCOLD_PSTEMP] = {&aslt, &bslt, &ultp1, &ultp2, &lgtp1, &lgtp2};
        for (int k = 0; k < NCOLD_PSTEMP; ++k) m[k]->p = c + k * LAND_TILE;
    }
    template <class SEG>
    __device__ __forceinline__ void refresh_daily(const SEG &s) {
        const int fb[6] = {FB_M_ASLT, FB_M_BSLT, FB_M_ULTP1, FB_M_ULTP2, FB_M_LGTP1, FB_M_LGTP2};
        const int mid[6] = {M_ASLT, M_BSLT, M_ULTP1, M_ULTP2, M_LGTP1, M_LGTP2};
        const int pid[6] = {P_ASLT, P_BSLT, P_ULTP1, P_ULTP2, P_LGTP1, P_LGTP2};
        double v[6];
        s.daily_batch(fb, mid, pid, v);
        aslt = v[0]; bslt = v[1]; ultp1 = v[2]; ultp2 = v[3]; lgtp1 = v[4]; lgtp2 = v[5];
    }
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
    const bool metric = s.metric();
    const double airtcs = (w.airts - 32.0) * 0.555;
    const double airtc = (airtmp - 32.0) * 0.555;
    if (hrfg) {
        const double aslt = metric ? (double)w.aslt : (w.aslt - 32.0) * 0.555;
        const double bslt = w.bslt;
        w.sltmp = aslt + bslt * airtc;
    }
    if (s.flag(FB_TSOP1)) {
        if (hrfg) {
            const double ault = metric ? (double)w.ultp1 : (w.ultp1 - 32.0) * 0.5555;
            const double bult = w.ultp2;
            w.ultmp = ault + bult * airtc;
            w.lgtmp = metric ? (double)w.lgtp1 : (w.lgtp1 - 32.0) * 0.5555;
        }
    } else {
        const double ulsmo = w.ultp1;
        const double ultdif = metric ? (double)w.ultp2 : 0.555 * w.ultp2;
        const double ultmps = w.ultmp;
        w.ultmp = ultmps + ulsmo * (airtcs + ultdif - ultmps);
        const double lgsmo = w.lgtp1;
        const double lgtdif = metric ? (double)w.lgtp2 : 0.555 * w.lgtp2;
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
    // metric: the series are stored in deg C; PWTGAS converts what it reads `(x * 9. / 5.) + 32.` (PWTGAS.py:168-170) -- the very
    // expression above, so `o` is what PWTGAS works on in both unit systems
    s.save(O_PSTEMP_AIRTC, metric ? airtc : (airtc * 9.0 / 5.0) + 32.0);
    s.save(O_PSTEMP_SLTMP, metric ? w.sltmp : o.sltmp);
    s.save(O_PSTEMP_ULTMP, metric ? w.ultmp : o.ultmp);
    s.save(O_PSTEMP_LGTMP, metric ? w.lgtmp : o.lgtmp);
}
