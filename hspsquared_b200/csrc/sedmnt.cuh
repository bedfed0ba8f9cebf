// sedmnt.cuh -- one SEDMNT interval (soil detachment / washoff / scour / re-attachment) for one segment.
//
// Behaviour: HSPF HPERSED as HSP2 runs it, src/hsp2/hsp2/SEDMNT.py:147-262 (_sedmnt_ loop body).  From the
// behavioural spec: VSIVFG is compared as a number (==2, <=2), the day-wet latch DRYDFG, and
// washoff "method 1" using surs+suro.
#pragma once
#include "land_common.cuh"

struct SedmntState {
    double dets, cover, nvsi;
    int drydfg;
    template <class SEG>
    __device__ __forceinline__ void load(const SEG &s) {
        dets = s.st(S_DETS); cover = s.st(S_SED_COVER); nvsi = s.st(S_SED_NVSI); drydfg = s.st(S_DRYDFG) != 0.0;
    }
    template <class SEG>
    __device__ __forceinline__ void store(const SEG &s) const {
        s.st(S_DETS, dets); s.st(S_SED_COVER, cover); s.st(S_SED_NVSI, nvsi); s.st(S_DRYDFG, drydfg ? 1.0 : 0.0);
    }
};

// rain: RAINF when SNOW (or the PWATER wrapper's NaN fill) put it in ts, else PREC (SEDMNT.py:86-89)
// metric runs: suro / surs arrive as the reference's `SURO * 0.0394` of the STORED series (land_kernels.cuh); rain is not
// converted by the reference; the stored series are converted by save() (land_common.cuh metric_out_code)
template <class SEG>
__device__ __forceinline__ void sedmnt_step(const SEG &s, SedmntState &w, double rain, double prec, double snocov,
                                            double suro, double surs) {
    const double delt = s.L.delt, delt60 = s.L.delt60;
    const bool dayfg = s.calfl & HSP2G_CAL_DAYFG;
    const double slsed = s.forcing(F_SLSED);
    if (dayfg) w.cover = s.daily(FB_M_COVER, M_COVER, P_COVER);

    // ---- DETACH (SEDMNT.py:161-172)
    if (rain > 0.0) {
        double cr;
        if (s.flag(FB_SED_CSNO1))
            cr = snocov > 0.0 ? w.cover + (1.0 - w.cover) * snocov : w.cover;
        else
            cr = w.cover;
        const double det = delt60 * (1.0 - cr) * s.par(P_SMPF) * s.par(P_KRER) * hsp_pow(rain / delt60, s.par(P_JRER));
        w.dets = w.dets + det;
    }
    if (dayfg) {  // SEDMNT.py:175-179
        w.nvsi = s.daily(FB_M_NVSI, M_NVSI, P_NVSI) * delt / 1440.;
        if (s.metric()) w.nvsi = HDIV(w.nvsi * 2.205, 2.471);  // SEDMNT.py:79,114: kg/ha to lbs/ac, after the time conversion
        if (s.flag(FB_VSIV_EQ2) && w.drydfg) {
            const double d = w.nvsi * (1440. / delt);
            w.dets = w.dets + d;
        }
    }
    double d = slsed;
    if (s.flag(FB_VSIV_LE2)) d = d + w.nvsi;
    w.dets = w.dets + d;

    double wssd, scrsd, sosed;
    if (suro > 0.0) {
        const double kser = s.par(P_KSER), jser = s.par(P_JSER), kger = s.par(P_KGER), jger = s.par(P_JGER);
        if (s.flag(FB_SDOPFG)) {  // method 1 (SEDMNT.py:188-212)
            const double arg = surs + suro;
            const double stcap = delt60 * kser * hsp_pow(arg / delt60, jser);
            if (stcap > w.dets)
                wssd = w.dets * suro / arg;
            else
                wssd = stcap * suro / arg;
            w.dets = w.dets - wssd;
            scrsd = delt60 * kger * hsp_pow(arg / delt60, jger);
            scrsd = scrsd * suro / arg;
            sosed = wssd + scrsd;
        } else {  // method 2 (SEDMNT.py:214-233)
            const double stcap = delt60 * kser * hsp_pow(suro / delt60, jser);
            if (stcap > w.dets) {
                wssd = w.dets;
                w.dets = 0.0;
            } else {
                wssd = stcap;
                w.dets = w.dets - wssd;
            }
            scrsd = delt60 * kger * hsp_pow(suro / delt60, jger);
            sosed = wssd + scrsd;
        }
    } else {
        wssd = 0.0;
        scrsd = 0.0;
        sosed = 0.0;
    }
    if (dayfg) {  // ATACH (SEDMNT.py:237-248)
        if (w.drydfg) w.dets = w.dets * (1.0 - s.par(P_AFFIX));
        w.drydfg = 1;
    }
    if (prec > 0) w.drydfg = 0;

    s.save(O_SEDMNT_DETS, w.dets);
    s.save(O_SEDMNT_WSSD, wssd);
    s.save(O_SEDMNT_SCRSD, scrsd);
    s.save(O_SEDMNT_SOSED, sosed);
    s.save(O_SEDMNT_COVER, w.cover);
}
