// iwater.cuh -- one IWATER interval (impervious-land water budget) for one segment.
//
// Behaviour: HSPF HIMPWAT as HSP2 runs it, src/hsp2/hsp2/IWATER.py:146-283 (_iwater_ loop body).  Written from
// the behavioural spec: note the 1.67 exponent of the ARM branch against 1.667 in the Newton
// branch, the unguarded dsuro/suro convergence test, and the petadj nesting that differs from PWATER.
#pragma once
#include "land_common.cuh"
#include "pwater.cuh"  // SnowLink

struct IwaterState {
    double rets, surs, msupy, dec, src, petadj;
    double retsc;  // daily-interpolated retention capacity

    template <class SEG>

    __device__ __forceinline__ void load(const SEG &s) {
        rets = s.st(S_RETS); surs = s.st(S_SURS); msupy = s.st(S_MSUPY); dec = s.st(S_DEC); src = s.st(S_SRC);
        petadj = s.st(S_PETADJ);
    }
    template <class SEG>
    __device__ __forceinline__ void store(const SEG &s) const {
        s.st(S_RETS, rets); s.st(S_SURS, surs); s.st(S_MSUPY, msupy); s.st(S_DEC, dec); s.st(S_SRC, src);
        s.st(S_PETADJ, petadj);
    }
    template <class SEG>
    __device__ __forceinline__ void refresh_daily(const SEG &s) {
        const double v = s.daily(FB_M_RETSC, M_RETSC, P_RETSC);
        retsc = s.metric() ? v * 0.0394 : v;  // IWATER.py:112-113: RETSC series * 0.0394 after the monthly interpolation
    }
};

template <bool HOURLY, class SEG>
__device__ __forceinline__ void iwater_step(const SEG &s, IwaterState &w, const SnowLink &sn, double prec, double &suro_out,
                                            double &surs_out) {
    const double delt60 = s.L.delt60;
    const bool hrfg = HOURLY ? true : (s.calfl & HSP2G_CAL_HRFG);
    const bool hr1fg = s.calfl & HSP2G_CAL_HR1FG;
    const double oldmsupy = w.msupy;
    const double petinp = s.forcing(F_PETINP);
    double supy, pet, petadj_out = 0.0;

    if (s.flag(FB_CSNOFG)) {  // IWATER.py:152-167
        supy = sn.rainf * (1.0 - sn.snocov) + sn.wyield;
        if (hrfg) {
            double pa = 1.0 - sn.snocov;
            if (sn.airtmp < s.par(P_PETMAX)) {
                if (sn.airtmp < s.par(P_PETMIN)) pa = 0.0;
                if (pa > 0.5) pa = 0.5;
            }
            w.petadj = pa;
            petadj_out = pa;
        }
        pet = petinp * w.petadj;
    } else {
        supy = prec;
        pet = petinp;
    }
    const double surli = s.forcing(F_SURLI);
    const double retsc = w.retsc;

    // ---- RETN (IWATER.py:175-196)
    const bool rtli = s.flag(FB_RTLIFG);
    const double reti = rtli ? supy + surli : supy;
    double reto;
    w.rets += reti;
    if (w.rets > retsc) {
        reto = w.rets - retsc;
        w.rets = retsc;
    } else
        reto = 0.0;
    const double suri = rtli ? reto : reto + surli;
    w.msupy = suri + w.surs;
    const double msupy = w.msupy;

    double suro = 0.0;
    int e0 = 0;
    if (msupy > 0.0002) {
        if (oldmsupy == 0.0 || hr1fg) {  // IWATER.py:204-207, 221-224
            const double slsur = s.par(P_SLSUR);
            const double dummy = s.daily(FB_M_NSUR, M_NSUR, P_NSUR) * s.par(P_LSUR);
            w.dec = 0.00982 * hsp_pow(dummy / sqrt(slsur), 0.6);
            w.src = 1020.0 * sqrt(slsur) / dummy;
        }
        if (s.flag(FB_RTOP1)) {  // ARM / NPS closed form (IWATER.py:202-218)
            const double sursm = (w.surs + msupy) * 0.5;
            double dummy = sursm * 1.6;
            if (suri > 0.0) {
                const double d = w.dec * hsp_pow(suri, 0.6);
                if (d > sursm) {
                    const double r = sursm / d;
                    dummy = sursm * (1.0 + 0.6 * (r * (r * r)));
                }
            }
            const double tsuro = delt60 * w.src * hsp_pow(dummy, 1.67);
            suro = tsuro > msupy ? msupy : tsuro;
            w.surs = tsuro > msupy ? 0.0 : msupy - suro;
        } else {  // Newton (IWATER.py:219-254)
            const double ssupr = suri / delt60;
            const double surse = ssupr > 0.0 ? w.dec * hsp_pow(ssupr, 0.6) : 0.0;
            double sursnw = msupy;
            suro = 0.0;
            bool converged = false;
            for (int count = 0; count < 100; ++count) {
                double ratio, fact;
                if (ssupr > 0.0) {
                    ratio = sursnw / surse;
                    fact = ratio <= 1.0 ? 1.0 + 0.6 * (ratio * (ratio * ratio)) : 1.6;
                } else {
                    fact = 1.6;
                    ratio = 1e30;
                }
                const double ffact = (delt60 * w.src * hsp_pow(fact, 1.667)) * hsp_pow(sursnw, 1.667);
                const double fsuro = ffact - suro;
                const double dfact = -1.667 * ffact;
                double dfsuro = dfact / sursnw - 1.0;
                if (ratio <= 1.0) dfsuro += (dfact / (fact * surse)) * 1.8 * (ratio * ratio);
                const double dsuro = fsuro / dfsuro;
                suro = suro - dsuro;
                sursnw = msupy - suro;
                if (fabs(dsuro / suro) < 0.01) {
                    converged = true;
                    break;
                }
            }
            if (!converged) e0 = 1;
            w.surs = sursnw;
        }
    } else {
        suro = msupy;
        w.surs = 0.0;
    }

    // ---- EVRETN (IWATER.py:259-269)
    double impev;
    if (w.rets > 0.0) {
        if (pet > w.rets) {
            impev = w.rets;
            w.rets = 0.0;
        } else {
            impev = pet;
            w.rets -= impev;
        }
    } else
        impev = 0.0;

    suro_out = suro;
    surs_out = w.surs;
    s.count(E_IWATER0, e0);
    s.save(O_IWATER_IMPEV, impev);
    s.save(O_IWATER_PET, pet);
    s.save(O_IWATER_PETADJ, petadj_out);
    s.save(O_IWATER_RETS, w.rets);
    s.save(O_IWATER_SUPY, supy);
    s.save(O_IWATER_SURI, suri);
    s.save(O_IWATER_SURO, suro);
    s.save(O_IWATER_SURS, w.surs);
}
