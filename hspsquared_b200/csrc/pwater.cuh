// pwater.cuh -- one PWATER interval (pervious-land water budget) for one segment, storages in registers.
//
// Behaviour: HSPF HPERWAT as HSP2 runs it, src/hsp2/hsp2/PWATER.py:263-688 (_pwater_ loop body) and :696-769
// (proute).  Written from the behavioural spec; expression order is the reference's.
// libm: pow / exp2 / exp / log are CUDA's (<= 2 ulp); sqrt and division are IEEE-exact.
#pragma once
#include "land_common.cuh"

struct PwaterState {
    double ceps, surs, uzs, ifws, lzs, agws;
    ColdD gwvs, msupy;  // touched once or twice per interval
    // change once a day / on the first wet interval / when LZS moves by 2 %: not in registers (land_common.cuh ColdD)
    ColdG dec, src;
    ColdD ifwk1, ifwk2, rlzrat, lzfrac, rparm, petadj;
    // interpolated-once-per-day parameters that every interval reads (refreshed at day starts)
    ColdD cepsc, uzsn, lzetp, intfw;

    // c: this lane's shared-memory scratch column; sb: this lane's column of the segment block.  The
    // two values that only wet intervals touch stay in their state rows in global memory (L2), the rest is shared.
    __device__ __forceinline__ void bind(double *c, double *sb) {
        ColdD *m[NCOLD_PWATER] = {&ifwk1, &ifwk2, &rlzrat, &lzfrac, &rparm, &petadj, &cepsc, &uzsn, &lzetp, &intfw, &gwvs, &msupy};
        for (int k = 0; k < NCOLD_PWATER; ++k) m[k]->p = c + k * LAND_TILE;
        dec.p = sb + (SB_STATE + S_DEC) * LAND_TILE;
        src.p = sb + (SB_STATE + S_SRC) * LAND_TILE;
    }

    template <class SEG>
    __device__ __forceinline__ void load(const SEG &s) {
        ceps = s.st(S_CEPS); surs = s.st(S_SURS); uzs = s.st(S_UZS); ifws = s.st(S_IFWS); lzs = s.st(S_LZS);
        agws = s.st(S_AGWS); gwvs = s.st(S_GWVS); msupy = s.st(S_MSUPY);
        ifwk1 = s.st(S_IFWK1); ifwk2 = s.st(S_IFWK2); rlzrat = s.st(S_RLZRAT); lzfrac = s.st(S_LZFRAC);
        rparm = s.st(S_RPARM); petadj = s.st(S_PETADJ);
    }
    template <class SEG>
    __device__ __forceinline__ void store(const SEG &s) const {
        s.st(S_CEPS, ceps); s.st(S_SURS, surs); s.st(S_UZS, uzs); s.st(S_IFWS, ifws); s.st(S_LZS, lzs);
        s.st(S_AGWS, agws); s.st(S_GWVS, gwvs); s.st(S_MSUPY, msupy);  // dec, src live in their state rows
        s.st(S_IFWK1, ifwk1); s.st(S_IFWK2, ifwk2); s.st(S_RLZRAT, rlzrat); s.st(S_LZFRAC, lzfrac);
        s.st(S_RPARM, rparm); s.st(S_PETADJ, petadj);
    }
    template <class SEG>
    __device__ __forceinline__ void refresh_daily(const SEG &s) {
        const int fb[4] = {FB_M_CEPSC, FB_M_UZSN, FB_M_LZETP, FB_M_INTFW};
        const int mid[4] = {M_CEPSC, M_UZSN, M_LZETP, M_INTFW};
        const int pid[4] = {P_CEPSC, P_UZSN, P_LZETP, P_INTFW};
        double v[4];
        s.daily_batch(fb, mid, pid, v);
        if (s.metric()) {  // CEPSC, UZSN series * 0.0394 after the monthly interpolation (PWATER.py:236-238)
            v[0] = v[0] * 0.0394;
            v[1] = v[1] * 0.0394;
        }
        cepsc = v[0];
        uzsn = v[1];
        lzetp = v[2];
        intfw = v[3];  // read on wet intervals only (PWATER.py:442), constant within a day like the others
    }
};

// what PWATER takes from SNOW / ATEMP of the same interval
struct SnowLink {
    double rainf, snocov, wyield, airtmp, packi;
};

// fluxes later sections consume
struct PwaterFlux {
    double suro, surs, ifwo, agwo;
};

static __device__ __constant__ double UZRA_TAB[10] = {0.0, 1.25, 1.50, 1.75, 2.00, 2.10, 2.20, 2.25, 2.5, 4.0};
static __device__ __constant__ double INTGRL_TAB[10] = {0.0, 1.29, 1.58, 1.92, 2.36, 2.81, 3.41, 3.8, 7.1, 3478.};

// numpy `argmax(x < table) - 1` (PWATER.py:388,395): first k with x < table[k], minus one; none (or NaN) -> -1
__device__ __forceinline__ int first_above_m1(double x, const double *tab) {
    int r = -1;
#pragma unroll
    for (int k = 9; k >= 0; --k)
        if (x < tab[k]) r = k - 1;
    return r;
}

// PROUTE (PWATER.py:696-769): overland flow routing.  Returns the Newton non-convergence count (0/1).
__device__ __forceinline__ int proute(double psur, bool rtop1, double delt60, double dec, double src,
                                      double &surs, double &suro) {
    int nonconv = 0;
    if (psur > 0.0002) {
        if (!rtop1) {
            const double ssupr = HDIV(psur - surs, delt60);
            const double surse = ssupr > 0.0 ? dec * hsp_pow(ssupr, 0.6) : 0.0;
            double sursnw = psur;
            suro = 0.0;
            bool converged = false;
            for (int count = 0; count < 100; ++count) {
                double ratio, fact;
                if (ssupr > 0.0) {
                    ratio = HDIV(sursnw, surse);
                    fact = ratio <= 1.0 ? 1.0 + 0.6 * (ratio * (ratio * ratio)) : 1.6;
                } else {
                    ratio = 1.0e30;
                    fact = 1.6;
                }
                const double ffact = (delt60 * src * hsp_pow(fact, 1.667)) * hsp_pow(sursnw, 1.667);
                const double fsuro = ffact - suro;
                const double dfact = -1.667 * ffact;
                double dfsuro = HDIV(dfact, sursnw) - 1.0;
                if (ratio <= 1.0) {
                    const double dterm = HDIV(dfact, fact * surse) * 1.8 * (ratio * ratio);
                    dfsuro = dfsuro + dterm;
                }
                const double dsuro = HDIV(fsuro, dfsuro);
                suro = suro - dsuro;
                if (suro <= 1.0e-10) suro = 0.0;
                sursnw = psur - suro;
                double change = 0.0;
                if (fabs(suro) > 0.0) change = fabs(HDIV(dsuro, suro));
                if (change < 0.01) {
                    converged = true;
                    break;
                }
            }
            if (!converged) nonconv = 1;
            surs = sursnw;
        } else {
            const double ssupr = psur - surs;
            const double sursm = (surs + psur) * 0.5;
            double dummy;
            if (ssupr > 0.0) {
                dummy = dec * hsp_pow(ssupr, 0.6);
                if (dummy > sursm) {
                    const double r = HDIV(sursm, dummy);
                    dummy = sursm * (1.0 + 0.6 * (r * (r * r)));
                } else
                    dummy = sursm * 1.6;
            } else
                dummy = sursm * 1.6;
            const double tsuro = delt60 * src * hsp_pow(dummy, 1.667);
            suro = tsuro > psur ? psur : tsuro;
            surs = tsuro > psur ? 0.0 : psur - suro;
        }
    } else {
        suro = psur;
        surs = 0.0;
    }
    if (suro <= 1.0e-10) suro = 0.0;
    return nonconv;
}

template <bool HOURLY, class SEG>
__device__ __forceinline__ void pwater_step(const SEG &s, PwaterState &w, const SnowLink &sn, double prec,
                                            PwaterFlux &fx) {
    const LandLaunch &L = s.L;
    const double delt60 = L.delt60;
    const bool dayfg = s.calfl & HSP2G_CAL_DAYFG;
    const bool hrfg = HOURLY ? true : (s.calfl & HSP2G_CAL_HRFG);
    const bool csnofg = s.flag(FB_CSNOFG);
    const double oldmsupy = w.msupy;
    const double lzetp = w.lzetp, cepsc = w.cepsc, uzsn = w.uzsn;
    const double infilt = s.par(P_INFILT_IVL);
    const double kvary = s.par(P_KVARY), lzsn = s.par(P_LZSN);
    const double petinp = s.forcing(F_PETINP);
    double inffac = 1.0;
    int e1 = 0, e2 = 0, e3 = 0, e6 = 0, e7 = 0, e8 = 0;

    // ---- supply and PET (PWATER.py:279-303)
    double supy, pet, petadj_out = 0.0;
    if (csnofg) {
        supy = sn.rainf * (1.0 - sn.snocov) + sn.wyield;
        if (hrfg) {
            const double forest = s.par(P_FOREST);
            double pa = (1.0 - forest) * (1.0 - sn.snocov) + forest;
            if ((sn.airtmp < s.par(P_PETMAX)) && (pa > 0.5)) pa = 0.5;
            if (sn.airtmp < s.par(P_PETMIN)) pa = 0.0;
            w.petadj = pa;
            petadj_out = pa;
        }
        pet = petinp * w.petadj;
        if (s.flag(FB_PW_ICEFG)) inffac = py_max(s.par(P_FZGL), 1.0 - s.par(P_FZG) * sn.packi);
    } else {
        supy = prec;
        pet = petinp;
    }
    // PWATER.py:302-303; with ICEFG off the reference reads an unassigned fzgl == 0.0 (oracle notes)
    if (s.flag(FB_IFFCFG2)) {
        const double fzgl = s.flag(FB_PW_ICEFG) ? s.par(P_FZGL) : 0.0;
        inffac = s.forcing(F_LGTMP) <= 0.0 ? fzgl : 1.0;
    }

    // ---- ICEPT (PWATER.py:305-312)
    w.ceps = w.ceps + supy + 0.0;
    double cepo = 0.0;
    if (w.ceps > cepsc) {
        cepo = w.ceps - cepsc;
        w.ceps = cepsc;
    }
    const double suri = cepo + s.forcing(F_SURLI);
    w.msupy = suri + w.surs + 0.0;
    const double msupy = w.msupy;
    double lzrat = HDIV(w.lzs, lzsn);
    double suro, ifwi, infil, uzi;

    if (msupy <= 0.0) {
        w.surs = 0.0;
        suro = 0.0;
        ifwi = 0.0;
        infil = 0.0;
        uzi = 0.0;
    } else {  // ---- SURFAC / DISPOS (PWATER.py:326-437)
        double ibar = HDIV(infilt, hsp_pow(lzrat, s.par(P_INFEXP)));
        if (inffac < 1.0) ibar = ibar * inffac;
        const double imax = ibar * s.par(P_INFILD);
        const double imin = ibar - (imax - ibar);
        if (dayfg || oldmsupy == 0.0) {
            const double slsur = s.par(P_SLSUR);
            const double dummy = s.daily(FB_M_NSUR, M_NSUR, P_NSUR) * s.par(P_LSUR);
            w.dec = 0.00982 * hsp_pow(HDIV(dummy, sqrt(slsur)), 0.6);
            w.src = 1020.0 * HDIV(sqrt(slsur), dummy);
        }
        const double ratio = py_max(1.0001, w.intfw * hsp_exp2(lzrat));
        double under, over;
        if (msupy <= imin) {
            under = msupy;
            over = 0.0;
        } else if (msupy > imax) {
            under = (imin + imax) * 0.5;
            over = msupy - under;
        } else {
            over = HDIV(((msupy - imin) * (msupy - imin)) * 0.5, imax - imin);
            under = msupy - over;
        }
        infil = under;
        if (over <= 0.0) {
            w.surs = 0.0;
            suro = 0.0;
            ifwi = 0.0;
            uzi = 0.0;
        } else {
            const double pdro = over;
            if (s.flag(FB_UZFG)) {  // UZINF2 (PWATER.py:366-379)
                const double uzrat = HDIV(w.uzs, uzsn);
                double uzfrac;
                if (uzrat < 2.0) {
                    const double k1 = 3.0 - uzrat;
                    uzfrac = 1.0 - (uzrat * 0.5) * hsp_pow(HDIV(1.0, 1.0 + k1), k1);
                } else {
                    const double k2 = (2.0 * uzrat) - 3.0;
                    uzfrac = hsp_pow(HDIV(1.0, 1.0 + k2), k2);
                }
                uzi = pdro * uzfrac;
            } else {  // UZINF table look-up (PWATER.py:381-404)
                const double uzraa = HDIV(w.uzs, uzsn);
                int kk = first_above_m1(uzraa, UZRA_TAB);
                if (kk == -1) {
                    kk = 8;
                    e1 = 1;
                }
                const double intga = INTGRL_TAB[kk] + HDIV((INTGRL_TAB[kk + 1] - INTGRL_TAB[kk]) * (uzraa - UZRA_TAB[kk]),
                                                           UZRA_TAB[kk + 1] - UZRA_TAB[kk]);
                const double intgb = HDIV(pdro, uzsn) + intga;
                kk = first_above_m1(intgb, INTGRL_TAB);
                if (kk == -1) {
                    e2 = 1;
                    kk = 8;
                }
                const double uzrab = UZRA_TAB[kk] + HDIV((UZRA_TAB[kk + 1] - UZRA_TAB[kk]) * (intgb - INTGRL_TAB[kk]),
                                                         INTGRL_TAB[kk + 1] - INTGRL_TAB[kk]);
                uzi = (uzrab - uzraa) * uzsn;
                if (uzi < -1.0e-3) e7 = 1;
                uzi = py_max(0.0, uzi);
            }
            if (uzi > pdro) uzi = pdro;
            const double uzfrac = HDIV(uzi, pdro);
            const double iimin = imin * ratio, iimax = imax * ratio;
            double over2;
            if (msupy <= iimin)
                over2 = 0.0;
            else if (msupy > iimax)
                over2 = msupy - (iimin + iimax) * 0.5;
            else
                over2 = HDIV(((msupy - iimin) * (msupy - iimin)) * 0.5, iimax - iimin);
            double psur = over2;
            const double pifwi = pdro - psur;
            ifwi = pifwi * (1.0 - uzfrac);
            if (psur <= 0.0) {
                w.surs = 0.0;
                suro = 0.0;
            } else {
                psur = psur * (1.0 - uzfrac);
                e6 = proute(psur, s.flag(FB_RTOP1), delt60, w.dec, w.src, w.surs, suro);
            }
        }
    }

    // ---- INTFLW (PWATER.py:440-455)
    if (dayfg) {
        const double kifw = HDIV(-hsp_log(s.daily(FB_M_IRC, M_IRC, P_IRC)), HDIV(24.0, delt60));
        w.ifwk2 = 1.0 - hsp_exp(-kifw);
        w.ifwk1 = 1.0 - HDIV(w.ifwk2, kifw);
    }
    const double inflo = ifwi + s.forcing(F_IFWLI);
    const double value = inflo + w.ifws;
    double ifwo;
    if (value > 0.00002) {
        ifwo = (w.ifwk1 * inflo) + (w.ifwk2 * w.ifws);
        w.ifws = value - ifwo;
    } else {
        ifwo = 0.0;
        w.ifws = 0.0;
        w.uzs = w.uzs + value;
    }

    // ---- UZONE (PWATER.py:457-468)
    double uzrat = HDIV(w.uzs, uzsn);
    w.uzs = w.uzs + uzi + s.forcing(F_UZLI) + 0.0;
    double perc = 0.0;
    if (uzrat - lzrat > 0.01) {
        const double d = uzrat - lzrat;
        perc = 0.1 * infilt * inffac * uzsn * (d * (d * d));
        if (perc > w.uzs) {
            perc = w.uzs;
            w.uzs = 0.0;
        } else
            w.uzs -= perc;
    }
    const double iperc = perc + infil + s.forcing(F_LZLI);

    // ---- LZONE (PWATER.py:473-487)
    const double lperc = iperc + 0.0;
    double lzi = 0.0;
    if (lperc > 0.0) {
        const bool ifrd = s.flag(FB_IFRDFG);
        if (fabs(lzrat - w.rlzrat) > 0.02 || ifrd) {
            w.rlzrat = lzrat;
            if (lzrat <= 1.0) {
                const double indx = 2.5 - 1.5 * lzrat;
                w.lzfrac = ifrd ? 1.0 : 1.0 - lzrat * hsp_pow(HDIV(1.0, 1.0 + indx), indx);
            } else {
                const double indx = 1.5 * lzrat - 0.5;
                w.lzfrac = ifrd ? hsp_exp(-1.0 * (lzrat - 1.0)) : hsp_pow(HDIV(1.0, 1.0 + indx), indx);
            }
        }
        lzi = w.lzfrac * lperc;
        w.lzs += lzi;
    }
    const double gwi = iperc + 0.0 - lzi;

    // ---- GWATER (PWATER.py:492-528)
    double igwi = 0.0, agwi = 0.0;
    if (gwi > 0.0) {
        igwi = s.par(P_DEEPFR) * gwi;
        agwi = gwi - igwi;
    }
    const double ainflo = agwi + s.forcing(F_AGWLI) + 0.0;
    double agwo = 0.0;
    const double kgw = s.par(P_KGW);
    if (kvary > 0.0) {
        w.gwvs += ainflo;
        if (dayfg) w.gwvs = w.gwvs > 0.0001 ? w.gwvs * 0.97 : 0.0;
        if (w.agws > 1.0e-20) {
            agwo = kgw * (1.0 + kvary * w.gwvs) * w.agws;
            const double avail = ainflo + w.agws;
            if (agwo > avail) {
                e3 = 1;
                agwo = avail;
            }
        }
    } else if (w.agws > 1.0e-20)
        agwo = kgw * w.agws;
    if (agwo < 1.0e-12) agwo = 0.0;
    w.agws = w.agws + (ainflo - agwo);
    if (w.agws < 0.0) {
        e8 = 1;
        w.agws = 0.0;
    }

    // ---- EVAPT (PWATER.py:539-626)
    double rempet = pet, taet = 0.0, baset = 0.0;
    const double basetp = s.par(P_BASETP);
    if (rempet > 0.0 && basetp > 0.0) {
        const double baspet = basetp * rempet;
        if (baspet > agwo) {
            baset = agwo;
            agwo = 0.0;
        } else {
            baset = baspet;
            agwo -= baset;
        }
        taet += baset;
        rempet -= baset;
    }
    double cepe = 0.0;
    if (rempet > 0.0 && w.ceps > 0.0) {
        if (rempet > w.ceps) {
            cepe = w.ceps;
            w.ceps = 0.0;
        } else {
            cepe = rempet;
            w.ceps -= cepe;
        }
        taet += cepe;
        rempet -= cepe;
    }
    double uzet = 0.0;
    if (rempet > 0.0) {
        if (w.uzs > 0.001) {
            uzrat = HDIV(w.uzs, uzsn);
            const double uzpet = uzrat > 2.0 ? rempet : 0.5 * uzrat * rempet;
            if (uzpet > w.uzs) {
                uzet = w.uzs;
                w.uzs = 0.0;
            } else {
                uzet = uzpet;
                w.uzs -= uzet;
            }
        }
        taet += uzet;
        rempet -= uzet;
    }
    double agwet = 0.0;
    const double agwetp = s.par(P_AGWETP);
    if (rempet > 0.0 && agwetp > 0.0) {
        const double gwpet = rempet * agwetp;
        if (gwpet > w.agws) {
            agwet = w.agws;
            w.agws = 0.0;
        } else {
            agwet = gwpet;
            w.agws -= agwet;
        }
        if (fabs(kvary) > 0.0) w.gwvs -= agwet;
        taet += agwet;
        rempet -= agwet;
    }
    if (dayfg) {
        lzrat = HDIV(w.lzs, lzsn);
        w.rparm = lzetp <= 0.99999 ? HDIV(HDIV(0.25, 1.0 - lzetp) * lzrat * delt60, 24.0) : 1.0e10;
    }
    double lzet = 0.0;
    if (rempet > 0.0 && w.lzs > 0.02) {
        double lzpet;
        if (lzetp >= 0.99999)
            lzpet = rempet * lzetp;
        else if (!s.flag(FB_VLE_GE2)) {
            lzpet = rempet > w.rparm ? 0.5 * w.rparm : rempet * (1.0 - HDIV(rempet, 2.0 * w.rparm));
            if (lzetp < 0.5) lzpet = lzpet * 2.0 * lzetp;
        } else
            lzpet = lzrat < 1.0 ? lzetp * lzrat * rempet : lzetp * rempet;
        lzet = lzpet < (w.lzs - 0.02) ? lzpet : w.lzs - 0.02;
        w.lzs -= lzet;
        taet += lzet;
        rempet -= lzet;
    }
    const double tgws = w.agws;

    fx.suro = suro;
    fx.surs = w.surs;
    fx.ifwo = ifwo;
    fx.agwo = agwo;

    if (e1 | e2 | e3 | e6 | e7 | e8) {  // rare: one test on the common path
        s.count(E_PWATER1, e1);
        s.count(E_PWATER2, e2);
        s.count(E_PWATER3, e3);
        s.count(E_PWATER6, e6);
        s.count(E_PWATER7, e7);
        s.count(E_PWATER8, e8);
    }

    s.save(O_PWATER_AGWET, agwet);
    s.save(O_PWATER_AGWI, agwi);
    s.save(O_PWATER_AGWO, agwo);
    s.save(O_PWATER_AGWS, w.agws);
    s.save(O_PWATER_BASET, baset);
    s.save(O_PWATER_CEPE, cepe);
    s.save(O_PWATER_CEPS, w.ceps);
    s.save(O_PWATER_GWVS, w.gwvs);
    s.save(O_PWATER_IFWI, ifwi);
    s.save(O_PWATER_IFWO, ifwo);
    s.save(O_PWATER_IFWS, w.ifws);
    s.save(O_PWATER_IGWI, igwi);
    s.save(O_PWATER_INFFAC, inffac);
    s.save(O_PWATER_INFIL, infil);
    s.save(O_PWATER_LZET, lzet);
    s.save(O_PWATER_LZI, lzi);
    s.save(O_PWATER_LZS, w.lzs);
    s.save(O_PWATER_PERC, perc);
    s.save(O_PWATER_PERO, suro + ifwo + agwo);
    s.save(O_PWATER_PERS, w.ceps + w.surs + w.ifws + w.uzs + w.lzs + tgws);
    s.save(O_PWATER_PET, pet);
    s.save(O_PWATER_PETADJ, petadj_out);
    s.save(O_PWATER_SUPY, supy);
    s.save(O_PWATER_SURI, suri);
    s.save(O_PWATER_SURO, suro);
    s.save(O_PWATER_SURS, w.surs);
    s.save(O_PWATER_TAET, taet);
    s.save(O_PWATER_TGWS, tgws);
    s.save(O_PWATER_UZET, uzet);
    s.save(O_PWATER_UZI, uzi);
    s.save(O_PWATER_UZS, w.uzs);
}
