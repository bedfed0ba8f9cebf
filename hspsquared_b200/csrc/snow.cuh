// snow.cuh -- one SNOW interval for one land segment, all pack state in registers.
//
// Behaviour follows HSPF's HPERSNO as HSP2 runs it: src/hsp2/hsp2/SNOW.py:235-585 (_snow_ loop body) and
// :589-598 (vapor).  Written from the behavioural spec; evaluation order and parenthesisation
// of every floating expression are the reference's, because parity is judged at 1e-9 and this file is compiled
// without FMA contraction.  The same code serves PERLND and IMPLND (configuration.py:48,51).
#pragma once
#include "land_common.cuh"

// COLD = ColdD (shared memory; any time step) or ColdW (write-through, hourly kernels): see land_common.cuh
template <class COLD>
struct SnowStateT {
    double packf, packi, packw, pdepth, snocov, neghts;
    // touched once or twice per interval (plus their SAVE store): shared memory
    ColdD covinx, dull, paktmp, rdenpf, skyclr, xlnmlt, oldprec;
    // recomputed on every hourly (HRFG) interval a pack exists: carried only between HRFG intervals of a sub-hourly
    // run and into the next task / chunk
    COLD snotmp, dewtmp, mneghs, packwc, compct, gmeltr, mostht, neght, rdnsn, snowep, vap, albedo;
    int hr6fg;

    static constexpr int SMEM_ROWS = NCOLD_SNOW2 + (sizeof(COLD) == sizeof(ColdD) ? NCOLD_SNOW : 0);

    // c: shared-memory scratch column (ColdD); sb: this lane's column of the segment block (ColdW)
    __device__ __forceinline__ void bind(double *c, double *sb) {
        COLD *m[NCOLD_SNOW] = {&snotmp, &dewtmp, &mneghs, &packwc, &compct, &gmeltr,
                               &mostht, &neght,  &rdnsn,  &snowep, &vap,    &albedo};
        const int sid[NCOLD_SNOW] = {S_SNOTMP, S_DEWTMP, S_MNEGHS, S_PACKWC, S_COMPCT, S_GMELTR,
                                     S_MOSTHT, S_NEGHT,  S_RDNSN,  S_SNOWEP, S_VAP,    S_ALBEDO};
        for (int k = 0; k < NCOLD_SNOW; ++k) m[k]->p = cold_home(m[k], c + k * LAND_TILE, sb + (SB_STATE + sid[k]) * LAND_TILE);
        ColdD *m2[NCOLD_SNOW2] = {&covinx, &dull, &paktmp, &rdenpf, &skyclr, &xlnmlt, &oldprec};
        for (int k = 0; k < NCOLD_SNOW2; ++k) m2[k]->p = c + (SMEM_ROWS - NCOLD_SNOW2 + k) * LAND_TILE;
    }
    static __device__ __forceinline__ double *cold_home(const ColdD *, double *smem, double *) { return smem; }
    static __device__ __forceinline__ double *cold_home(const ColdW *, double *, double *glob) { return glob; }

    template <class SEG>
    __device__ __forceinline__ void load(const SEG &s) {
        covinx = s.st(S_COVINX); dull = s.st(S_DULL); packf = s.st(S_PACKF); packi = s.st(S_PACKI);
        packw = s.st(S_PACKW); paktmp = s.st(S_PAKTMP); rdenpf = s.st(S_RDENPF); skyclr = s.st(S_SKYCLR);
        xlnmlt = s.st(S_XLNMLT); pdepth = s.st(S_PDEPTH); snocov = s.st(S_SNOCOV); neghts = s.st(S_NEGHTS);
        oldprec = s.st(S_OLDPREC); hr6fg = s.st(S_HR6FG) != 0.0;
        load_cold(snotmp, s, S_SNOTMP); load_cold(dewtmp, s, S_DEWTMP); load_cold(mneghs, s, S_MNEGHS);
        load_cold(packwc, s, S_PACKWC); load_cold(compct, s, S_COMPCT); load_cold(gmeltr, s, S_GMELTR);
        load_cold(mostht, s, S_MOSTHT); load_cold(neght, s, S_NEGHT); load_cold(rdnsn, s, S_RDNSN);
        load_cold(snowep, s, S_SNOWEP); load_cold(vap, s, S_VAP); load_cold(albedo, s, S_ALBEDO);
    }
    template <class SEG>
    static __device__ __forceinline__ void load_cold(ColdD &d, const SEG &s, int sid) { d = s.st(sid); }
    template <class SEG>
    static __device__ __forceinline__ void load_cold(ColdW &d, const SEG &s, int sid) { d.v = s.st(sid); }  // no write-back

    template <class SEG>
    __device__ __forceinline__ void store(const SEG &s) const {
        s.st(S_COVINX, covinx); s.st(S_DULL, dull); s.st(S_PACKF, packf); s.st(S_PACKI, packi);
        s.st(S_PACKW, packw); s.st(S_PAKTMP, paktmp); s.st(S_RDENPF, rdenpf); s.st(S_SKYCLR, skyclr);
        s.st(S_XLNMLT, xlnmlt); s.st(S_PDEPTH, pdepth); s.st(S_SNOCOV, snocov); s.st(S_NEGHTS, neghts);
        s.st(S_OLDPREC, oldprec); s.st(S_HR6FG, hr6fg ? 1.0 : 0.0);
        store_cold(snotmp, s, S_SNOTMP); store_cold(dewtmp, s, S_DEWTMP); store_cold(mneghs, s, S_MNEGHS);
        store_cold(packwc, s, S_PACKWC); store_cold(compct, s, S_COMPCT); store_cold(gmeltr, s, S_GMELTR);
        store_cold(mostht, s, S_MOSTHT); store_cold(neght, s, S_NEGHT); store_cold(rdnsn, s, S_RDNSN);
        store_cold(snowep, s, S_SNOWEP); store_cold(vap, s, S_VAP); store_cold(albedo, s, S_ALBEDO);
    }
    template <class SEG>
    static __device__ __forceinline__ void store_cold(const ColdD &d, const SEG &s, int sid) { s.st(sid, d); }
    template <class SEG>
    static __device__ __forceinline__ void store_cold(const ColdW &, const SEG &, int) {}  // already written through
};

// fluxes of the current interval that the water sections consume
struct SnowFlux {
    double rainf, snowf, prain, snowe, melt, wyield;
};

// saturation vapour pressure table, utilities.py:52-55
static __device__ __constant__ double SVP_TAB[40] = {1.005, 1.005, 1.005, 1.005, 1.005, 1.005, 1.005, 1.005, 1.005, 1.005,
                                              1.01,  1.01,  1.015, 1.02,  1.03,  1.04,  1.06,  1.08,  1.1,   1.29,
                                              1.66,  2.13,  2.74,  3.49,  4.40,  5.55,  6.87,  8.36,  10.1,  12.2,
                                              14.6,  17.5,  20.9,  24.8,  29.3,  34.6,  40.7,  47.7,  55.7,  64.9};

// SNOW.py:589-598.  A NaN temperature lands "below the table" exactly like Numba's int(floor(nan)).
__device__ __forceinline__ double svp_at(double temp) {
    const double indx = (temp + 100.0) * 0.2 - 1.0;
    const double fl = floor(indx);
    if (!(fl >= 0.0)) return 1.005;
    if (fl + 1.0 > 39.0) return 64.9;
    const int lower = (int)fl;
    const double a = SVP_TAB[lower], b = SVP_TAB[lower + 1];
    return a + (indx - (double)lower) * (b - a);
}

template <bool HOURLY, class SEG, class SNOWSTATE>
__device__ __forceinline__ void snow_step(const SEG &s, SNOWSTATE &w, SnowFlux &x, double airtmp, double prec) {
    const LandLaunch &L = s.L;
    const double delt = L.delt, delt60 = L.delt60;
    const bool hrfg = HOURLY ? true : (s.calfl & HSP2G_CAL_HRFG);
    const bool snop = s.flag(FB_SNOPFG);
    const double tsnow = s.par(P_TSNOW);
    const double reltmp = airtmp - 32.0;
    const double oldprec = w.oldprec;
    w.oldprec = prec;

    // ---- METEOR (SNOW.py:255-290): rain/snow split, dewpoint, sky clearness
    const bool fprfg = (prec > 0.0) ? (oldprec == 0.0) : false;
    if (hrfg) {
        const double dtmpg = s.forcing(F_DTMPG);
        w.dewtmp = ((prec > 0.0 && airtmp > tsnow) || dtmpg > airtmp) ? airtmp : dtmpg;
    }
    double rainf, snowf;
    if (prec > 0.0) {
        if (hrfg || fprfg) {
            const double dtsnow = (airtmp - w.dewtmp) * (0.12 + 0.008 * airtmp);
            w.snotmp = tsnow + py_min(1.0, dtsnow);
        }
        if (!snop) w.skyclr = 0.15;
        if (airtmp < w.snotmp) {
            snowf = prec * s.par(P_SNOWCF);
            rainf = 0.0;
            if (hrfg || fprfg) {
                const double rdcsn = s.par(P_RDCSN);
                const double a = hsp_div100(airtmp);
                w.rdnsn = airtmp > 0.0 ? rdcsn + a * a : rdcsn;
            }
        } else {
            rainf = prec;
            snowf = 0.0;
        }
    } else {
        rainf = 0.0;
        snowf = 0.0;
        if (!snop && w.skyclr < 1.0) {
            w.skyclr += (0.0004 * delt);
            if (w.skyclr > 1.0) w.skyclr = 1.0;
        }
    }
    if (!snop && s.has_forcing(F_CLOUD)) w.skyclr = py_max(0.15, 1.0 - hsp_div10(s.forcing(F_CLOUD)));

    double prain, snowe, melt, wyield;
    if (w.packf > 0.0 || snowf > 0.0) {
        bool iregfg;
        if (w.packf == 0.0) {
            iregfg = true;
            if (!snop) w.dull = 0.0;
        } else
            iregfg = hrfg;
        const double covind = s.par(P_COVIND);

        // ---- EFFPRC (SNOW.py:302-317)
        if (snowf > 0.0) {
            w.packf += snowf;
            w.pdepth += HDIV(snowf, w.rdnsn);
            if (w.packf > w.covinx) w.covinx = w.packf > covind ? covind : w.packf;
            if (!snop) {
                const double d = 1000.0 * snowf;
                w.dull = d >= w.dull ? 0.0 : w.dull - d;
            }
            prain = 0.0;
        } else
            prain = rainf > 0.0 ? rainf * w.snocov : 0.0;
        if (!snop) {
            if (w.dull < 800.0) w.dull += delt60;
        }

        // ---- COMPAC (SNOW.py:319-326)
        if (iregfg) {
            w.rdenpf = HDIV(w.packf, w.pdepth);
            const double d = 1.0 - (0.00002 * delt60 * w.pdepth * (0.55 - w.rdenpf));
            w.compct = w.rdenpf < 0.55 ? d : 1.0;
        }
        if (w.compct < 1.0) w.pdepth *= w.compct;

        // ---- SNOWEV (SNOW.py:328-349)
        const double winmov = s.forcing(F_WINMOV);
        if (!snop) {
            if (iregfg) {
                w.vap = svp_at(w.dewtmp);
                const double satvap = svp_at(airtmp);
                const double d = s.par(P_SNOEVP_K) * winmov * (satvap - w.vap) * w.snocov;  // SNOEVP_K = SNOEVP * 0.0002
                w.snowep = w.vap >= 6.108 ? 0.0 : d;
            }
            if (w.snowep >= w.packf) {
                snowe = w.packf;
                w.pdepth = 0.0;
                w.packi = 0.0;
                w.packf = 0.0;
            } else {
                w.pdepth *= (1.0 - HDIV(w.snowep, w.packf));
                w.packf -= w.snowep;
                snowe = w.snowep;
                if (w.packi > w.packf) w.packi = w.packf;
            }
        } else
            snowe = 0.0;

        if (iregfg) {
            if (!snop) {  // ---- HEXCHR + ALBEDO (SNOW.py:353-377)
                const double factr = s.par(P_CCFACT_K) * winmov;  // CCFACT_K = CCFACT * 0.00026
                double d = 8.59 * (w.vap - 6.108);
                const double condht = w.vap > 6.108 ? d * factr : 0.0;
                d = reltmp * s.par(P_MELEV_FAC) * factr;  // MELEV_FAC = 1.0 - 0.3 * MELEV / 10000.0 (host, segstore.py)
                const double convht = airtmp > 32.0 ? d : 0.0;
                d = sqrt(hsp_div24(w.dull));
                const double summer = py_max(0.45, 0.80 - 0.10 * d);
                const double winter = py_max(0.60, 0.85 - 0.07 * d);
                w.albedo = (s.calfl & HSP2G_CAL_SEASON) ? summer : winter;
                const double shade = s.par(P_SHADE);
                const double kk = (1.0 - shade) * delt60;
                const double long1 = (shade * 0.26 * reltmp) + kk * (0.20 * reltmp - 6.6);
                const double long2 = (shade * 0.20 * reltmp) + kk * (0.17 * reltmp - 6.6);
                double lng = reltmp > 0.0 ? long1 : long2;
                if (lng < 0.0) lng *= w.skyclr;
                const double shrt = s.forcing(F_SOLRAD) * (1.0 - w.albedo) * (1.0 - shade);
                w.mostht = HDIV(shrt + lng, 203.2) + convht + condht;
            } else  // ---- DEGDAY (SNOW.py:380); KMELT*delt/1440 as in SNOW.py:145
                w.mostht = (s.daily(FB_M_KMELT, M_KMELT, P_KMELT) * delt / 1440.0) * (airtmp - s.par(P_TBASE));
        }
        const double rnsht = rainf > 0.0 ? hsp_div144(reltmp * rainf) : 0.0;
        double sumht = w.mostht + rnsht;
        if (w.snocov < 1.0) sumht = sumht * w.snocov;
        w.paktmp = w.neghts <= 0.0 ? 32.0 : 32.0 - HDIV(w.neghts, 0.00695 * w.packf);

        // ---- COOLER (SNOW.py:391-399)
        if (iregfg) {
            w.mneghs = reltmp > 0.0 ? 0.0 : -reltmp * 0.00695 * w.packf / 2.0;
            w.neght = w.paktmp <= airtmp ? 0.0 : 0.0007 * (w.paktmp - airtmp) * delt60;
        }
        if (sumht < 0.0) {
            if (w.paktmp > airtmp) w.neghts = py_min(w.mneghs, w.neghts + w.neght);
            sumht = 0.0;
        }

        // ---- WARMUP (SNOW.py:402-426)
        double rnfrz = 0.0;
        if (w.neghts > 0.0) {
            if (sumht > 0.0) {
                if (sumht > w.neghts) {
                    sumht -= w.neghts;
                    w.neghts = 0.0;
                } else {
                    w.neghts -= sumht;
                    sumht = 0.0;
                }
            }
            if (prain > 0.0) {
                if (prain > w.neghts) {
                    rnfrz = w.neghts;
                    w.packf += rnfrz;
                    w.neghts = 0.0;
                } else {
                    rnfrz = prain;
                    w.neghts -= prain;
                    w.packf += prain;
                }
                if (w.packf > w.pdepth) w.pdepth = w.packf;
            }
        }

        // ---- MELTER (SNOW.py:428-441)
        if (sumht >= w.packf) {
            melt = w.packf;
            w.packf = 0.0;
            w.pdepth = 0.0;
            w.packi = 0.0;
        } else if (sumht > 0.0) {
            melt = sumht;
            w.pdepth *= (1.0 - HDIV(melt, w.packf));
            w.packf -= melt;
            if (w.packi > w.packf) w.packi = w.packf;
        } else
            melt = 0.0;

        // ---- LIQUID (SNOW.py:443-458)
        if (iregfg && w.packf > 0.0) {
            const double mwater = s.par(P_MWATER);
            w.rdenpf = HDIV(w.packf, w.pdepth);
            if (w.rdenpf <= 0.6)
                w.packwc = mwater;
            else {
                const double df = 3.0 - 3.33 * w.rdenpf;
                w.packwc = df >= 0.0 ? mwater * df : 0.0;
            }
        }
        const double pwsupy = w.packw + melt + prain - rnfrz;
        const double mpws = w.packwc * w.packf;
        if ((pwsupy - mpws) > (0.01 * delt60)) {
            wyield = pwsupy - mpws;
            w.packw = mpws;
        } else {
            w.packw = pwsupy;
            wyield = 0.0;
        }

        // ---- ICING (SNOW.py:461-485)
        if (s.flag(FB_ICEFG)) {
            if (!(s.calfl & HSP2G_CAL_HR6IND)) {
                if (w.hr6fg) {
                    if (w.snocov < 1.0) {
                        const double xlnem = -reltmp * 0.01;
                        if (xlnem > w.xlnmlt) w.xlnmlt = xlnem;
                    }
                    w.hr6fg = 0;
                }
            } else
                w.hr6fg = 1;
            if (wyield > 0.0 && w.xlnmlt > 0.0) {
                double freeze;
                if (wyield < w.xlnmlt) {
                    freeze = wyield;
                    w.xlnmlt -= wyield;
                    wyield = 0.0;
                } else {
                    freeze = w.xlnmlt;
                    wyield -= w.xlnmlt;
                    w.xlnmlt = 0.0;
                }
                w.packf += freeze;
                w.packi += freeze;
                w.pdepth += freeze;
            }
        }

        // ---- GMELT (SNOW.py:487-508)
        if (iregfg) {
            const double mgmelt = s.par(P_MGMELT_IVL);
            if (w.paktmp >= 32.0)
                w.gmeltr = mgmelt;
            else if (w.paktmp > 5.0)
                w.gmeltr = mgmelt * (1.0 - 0.03 * (32.0 - w.paktmp));
            else
                w.gmeltr = mgmelt * 0.19;
        }
        if (w.packf <= w.gmeltr) {
            wyield = wyield + w.packf + w.packw;
            w.packf = 0.0;
            w.packi = 0.0;
            w.packw = 0.0;
            w.pdepth = 0.0;
            w.neghts = 0.0;
        } else {
            const double d = 1.0 - HDIV(w.gmeltr, w.packf);
            w.packw += w.gmeltr;
            w.pdepth *= d;
            w.neghts *= d;
            w.packf -= w.gmeltr;
            w.packi = w.packi > w.gmeltr ? w.packi - w.gmeltr : 0.0;
        }

        // ---- pack bookkeeping / pack vanishes (SNOW.py:511-537)
        if (w.packf > 0.005) {
            w.rdenpf = HDIV(w.packf, w.pdepth);
            w.paktmp = w.neghts == 0.0 ? 32.0 : 32.0 - HDIV(w.neghts, 0.00695 * w.packf);
            w.snocov = w.packf < w.covinx ? HDIV(w.packf, w.covinx) : 1.0;
        } else {
            melt += w.packf;
            wyield += w.packf + w.packw;
            w.hr6fg = 1;
            w.packf = 0.0;
            w.packi = 0.0;
            w.packw = 0.0;
            w.pdepth = 0.0;
            w.rdenpf = __longlong_as_double(0x7ff8000000000000LL);
            w.covinx = 0.1 * covind;
            w.snocov = 0.0;
            w.xlnmlt = 0.0;
            w.mneghs = __longlong_as_double(0x7ff8000000000000LL);
            w.paktmp = 32.0;
            w.neghts = 0.0;
            w.dull = -1.0E30;
            w.albedo = -1.0E30;
            w.snowep = -1.0E30;
            w.vap = -1.0E30;
        }
    } else {  // SNOW.py:538-543
        prain = 0.0;
        snowe = 0.0;
        wyield = 0.0;
        melt = 0.0;
    }
    x.rainf = rainf;
    x.snowf = snowf;
    x.prain = prain;
    x.snowe = snowe;
    x.melt = melt;
    x.wyield = wyield;

    s.save(O_SNOW_ALBEDO, w.albedo);
    s.save(O_SNOW_COVINX, w.covinx);
    s.save(O_SNOW_DEWTMP, w.dewtmp);
    s.save(O_SNOW_DULL, w.dull);
    s.save(O_SNOW_MELT, melt);
    s.save(O_SNOW_NEGHTS, w.neghts);
    s.save(O_SNOW_PACKF, w.packf);
    s.save(O_SNOW_PACKI, w.packi);
    s.save(O_SNOW_PACKW, w.packw);
    // metric: PACK[step] = PACKF[step] + PACKW[step] of the CONVERTED series (SNOW.py:576); save() leaves PACK itself as is
    s.save(O_SNOW_PACK, s.metric() ? (w.packf * 25.4) + (w.packw * 25.4) : w.packf + w.packw);
    s.save(O_SNOW_PAKTMP, w.paktmp);
    s.save(O_SNOW_PDEPTH, w.pdepth);
    s.save(O_SNOW_PRAIN, prain);
    s.save(O_SNOW_RAINF, rainf);
    s.save(O_SNOW_RDENPF, w.rdenpf);
    s.save(O_SNOW_SKYCLR, w.skyclr);
    s.save(O_SNOW_SNOCOV, w.snocov);
    s.save(O_SNOW_SNOTMP, w.snotmp);
    s.save(O_SNOW_SNOWE, snowe);
    s.save(O_SNOW_SNOWF, snowf);
    s.save(O_SNOW_WYIELD, wyield);
    s.save(O_SNOW_XLNMLT, w.xlnmlt);
}
