/*
 * hsp2_oracle.c -- CPU ORACLE (test infrastructure, not product).  See hsp2_oracle.h for scope and pinning.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -fopenmp  (oracle/Makefile).
 * The reference's Numba code compiles without FMA contraction or fast-math and reaches glibc libm for
 * pow/exp/exp2/log/sqrt, so with these flags this file is bit-identical to it.
 * Python `max(a,b)` / `min(a,b)` are `b if b>a else a` / `b if b<a else a` (NaN semantics differ from fmax/fmin).
 */
#include "hsp2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline double py_max(double a, double b) { return b > a ? b : a; }
static inline double py_min(double a, double b) { return b < a ? b : a; }

#define STR_(x) #x
#define HSP2O_NAME(n) STR_(n) ","
#define FIELDS(LIST) LIST(HSP2O_NAME)

const char *hsp2o_fields(const char *r) {
#define REC(name, UI, IN, OUT)                                  \
    if (!strcmp(r, #name "_ui")) return FIELDS(UI);              \
    if (!strcmp(r, #name "_in")) return FIELDS(IN);              \
    if (!strcmp(r, #name "_out")) return FIELDS(OUT);
    REC(atemp, ATEMP_UI, ATEMP_IN, ATEMP_OUT)
    REC(snow, SNOW_UI, SNOW_IN, SNOW_OUT)
    REC(pwater, PWATER_UI, PWATER_IN, PWATER_OUT)
    REC(iwater, IWATER_UI, IWATER_IN, IWATER_OUT)
    REC(sedmnt, SEDMNT_UI, SEDMNT_IN, SEDMNT_OUT)
    REC(pstemp, PSTEMP_UI, PSTEMP_IN, PSTEMP_OUT)
    REC(pwtgas, PWTGAS_UI, PWTGAS_IN, PWTGAS_OUT)
    REC(solids, SOLIDS_UI, SOLIDS_IN, SOLIDS_OUT)
    REC(iwtgas, IWTGAS_UI, IWTGAS_IN, IWTGAS_OUT)
    REC(pqual, PQUAL_UI, PQUAL_IN, PQUAL_OUT)
    REC(iqual, IQUAL_UI, IQUAL_IN, IQUAL_OUT)
#undef REC
    if (!strcmp(r, "batch_par")) return FIELDS(BATCH_PAR);
    return "";
}

/* ============================================================================================== */
/* ATEMP  (ATEMP.py:33-59): wet lapse rate 0.0035 when PREC > k, else hour-of-day table            */
/* ============================================================================================== */
void hsp2o_atemp(const atemp_ui *ui, const atemp_in *in, const atemp_out *out, int64_t nsteps) {
    const double k = ui->k, eldat = ui->ELDAT;
    for (int64_t t = 0; t < nsteps; ++t) {
        double lapse = in->PREC[t] > k ? 0.0035 : in->LAPSE[t];
        out->AIRTMP[t] = in->GATMP[t] - lapse * eldat;
    }
}

/* ============================================================================================== */
/* SNOW  (SNOW.py:91-598)                                                                          */
/* ============================================================================================== */
/* saturation vapour pressure table, utilities.py:52-55 */
static const double SVP_TABLE[40] = {1.005, 1.005, 1.005, 1.005, 1.005, 1.005, 1.005, 1.005, 1.005, 1.005,
                                     1.01,  1.01,  1.015, 1.02,  1.03,  1.04,  1.06,  1.08,  1.1,   1.29,
                                     1.66,  2.13,  2.74,  3.49,  4.40,  5.55,  6.87,  8.36,  10.1,  12.2,
                                     14.6,  17.5,  20.9,  24.8,  29.3,  34.6,  40.7,  47.7,  55.7,  64.9};

/* SNOW.py:589-598; a NaN temperature makes int(floor(nan)) = INT64_MIN in Numba => "below table" */
static double svp_lookup(double temp) {
    double indx = (temp + 100.0) * 0.2 - 1.0;
    double fl = floor(indx);
    if (!(fl >= 0.0)) return 1.005; /* lower < 0 (or NaN) */
    if (fl + 1.0 > 39.0) return 64.9;
    int lower = (int)fl;
    return SVP_TABLE[lower] + (indx - (double)lower) * (SVP_TABLE[lower + 1] - SVP_TABLE[lower]);
}

void hsp2o_snow(const snow_ui *ui, const snow_in *in, const snow_out *out, int64_t nsteps, int64_t *errors) {
    if (errors) errors[0] = 0;
    const double delt = ui->delt;
    const double delt60 = delt / 60.0;
    const int cloudfg = (int)ui->cloudfg, icefg = (int)ui->ICEFG, snopfg = (int)ui->SNOPFG;
    double melev = ui->MELEV, tbase = ui->TBASE, tsnow = ui->TSNOW;
    const double rdcsn = ui->RDCSN;
    const int metric = ui->uunits == 2.0; /* SNOW.py:99 */

    /* SNOW.py:102-135 initial named states; PACKF includes the ice at load */
    double covinx = ui->COVINX, dull = ui->DULL;
    double packf = ui->PACKF + ui->PACKI, packi = ui->PACKI, packw = ui->PACKW, paktmp = ui->PAKTMP;
    double rdenpf = ui->RDENPF, skyclr = ui->SKYCLR, xlnmlt = ui->XLNMLT;
    double pdepth, snocov, neghts;
    if (metric) { /* SNOW.py:114-120, 131-133: `x * 9./5.` is (x * 9.) / 5. */
        packf = packf / 25.4;
        packi = packi / 25.4;
        packw = packw / 25.4;
        covinx = covinx / 25.4;
        paktmp = (paktmp * 9. / 5.) + 32.;
        melev = melev * 3.28084;
        tbase = (tbase * 9. / 5.) + 32.;
        tsnow = (tsnow * 9. / 5.) + 32.;
    }
    const double covind0 = metric ? in->COVIND[0] / 25.4 : in->COVIND[0]; /* SNOW.py:157-159 before :185-206 */

    if (packf + packw <= 1.0e-5) { /* SNOW.py:185-200 no pack at start */
        covinx = 0.1 * covind0;
        dull = 0.0;
        neghts = 0.0;
        packf = packi = packw = 0.0;
        paktmp = 32.0;
        pdepth = 0.0;
        rdenpf = NAN;
        snocov = 0.0;
    } else { /* SNOW.py:201-206 */
        if (covinx < 1.0e-5) covinx = 0.1 * covind0;
        pdepth = packf / rdenpf;
        snocov = packf < covinx ? packf / covinx : 1.0;
        neghts = (32.0 - paktmp) * 0.00695 * packf;
    }
    /* SNOW.py:208-227 */
    double melt = 0.0, mneghs = 0.0, packwc = 0.0, prain = 0.0, prec = 0.0, snotmp = tsnow, snowe = 0.0,
           wyield = 0.0;
    double albedo = 0.0, compct = 0.0, dewtmp = 0.0, gmeltr = 0.0, mostht = 0.0, neght = 0.0, rdnsn = 0.0,
           satvap = 0.0, snowep = 0.0, vap = 0.0;
    int hr6fg = ((int64_t)in->HR6IND[0]) > 0 ? 1 : 0;

    for (int64_t t = 0; t < nsteps; ++t) {
        const double oldprec = prec;
        double mgmelt = in->MGMELT[t] * delt / 1440.0; /* SNOW.py:146 */
        const double airtmp = in->AIRTMP[t];
        double covind = in->COVIND[t];
        if (metric) { /* SNOW.py:157-159 */
            covind = covind / 25.4;
            mgmelt = mgmelt / 25.4;
        }
        const double dtmpg = in->DTMPG[t];
        const int hr6ind = (int)(int64_t)in->HR6IND[t];
        const int hrfg = (int)(int64_t)in->HRFG[t];
        const double mwater = in->MWATER[t];
        prec = in->PREC[t];
        const double shade = in->SHADE[t];
        const double snowcf = in->SNOWCF[t];
        const double solrad = in->SOLRAD[t];
        const double winmov = in->WINMOV[t];
        const double reltmp = airtmp - 32.0;
        double rainf, snowf;

        /* --- METEOR (SNOW.py:255-290) --- */
        const int fprfg = (prec > 0.0) ? (oldprec == 0.0) : 0;
        if (hrfg) dewtmp = ((prec > 0.0 && airtmp > tsnow) || dtmpg > airtmp) ? airtmp : dtmpg;
        if (prec > 0.0) {
            if (hrfg || fprfg) {
                double dtsnow = (airtmp - dewtmp) * (0.12 + 0.008 * airtmp);
                snotmp = tsnow + py_min(1.0, dtsnow);
            }
            if (snopfg == 0) skyclr = 0.15;
            if (airtmp < snotmp) {
                snowf = prec * snowcf;
                rainf = 0.0;
                if (hrfg || fprfg) rdnsn = airtmp > 0.0 ? rdcsn + (airtmp / 100.0) * (airtmp / 100.0) : rdcsn;
            } else {
                rainf = prec;
                snowf = 0.0;
            }
        } else {
            rainf = 0.0;
            snowf = 0.0;
            if (snopfg == 0 && skyclr < 1.0) {
                skyclr += (0.0004 * delt);
                if (skyclr > 1.0) skyclr = 1.0;
            }
        }
        if (snopfg == 0 && cloudfg) skyclr = py_max(0.15, 1.0 - (in->CLOUD[t] / 10.0));

        if (packf > 0.0 || snowf > 0.0) {
            int iregfg;
            if (packf == 0.0) {
                iregfg = 1;
                if (snopfg == 0) dull = 0.0;
            } else
                iregfg = hrfg;

            /* --- EFFPRC (SNOW.py:302-317) --- */
            if (snowf > 0.0) {
                packf += snowf;
                pdepth += (snowf / rdnsn);
                if (packf > covinx) covinx = packf > covind ? covind : packf;
                if (snopfg == 0) {
                    double d = 1000.0 * snowf;
                    dull = d >= dull ? 0.0 : dull - d;
                }
                prain = 0.0;
            } else
                prain = rainf > 0.0 ? rainf * snocov : 0.0;
            if (snopfg == 0) {
                if (dull < 800) dull += delt60;
            }

            /* --- COMPAC (SNOW.py:319-326) --- */
            if (iregfg) {
                rdenpf = packf / pdepth;
                double d = 1.0 - (0.00002 * delt60 * pdepth * (0.55 - rdenpf));
                compct = rdenpf < 0.55 ? d : 1.0;
            }
            if (compct < 1.0) pdepth *= compct;

            /* --- SNOWEV (SNOW.py:328-349) --- */
            if (snopfg == 0) {
                if (iregfg) {
                    vap = svp_lookup(dewtmp);
                    satvap = svp_lookup(airtmp);
                    double d = in->SNOEVP[t] * 0.0002 * winmov * (satvap - vap) * snocov;
                    snowep = vap >= 6.108 ? 0.0 : d;
                }
                if (snowep >= packf) {
                    snowe = packf;
                    pdepth = 0.0;
                    packi = 0.0;
                    packf = 0.0;
                } else {
                    pdepth *= (1.0 - snowep / packf);
                    packf -= snowep;
                    snowe = snowep;
                    if (packi > packf) packi = packf;
                }
            } else
                snowe = 0.0;

            if (iregfg) {
                if (snopfg == 0) { /* --- HEXCHR + ALBEDO (SNOW.py:353-377) --- */
                    double factr = in->CCFACT[t] * 0.00026 * winmov;
                    double d = 8.59 * (vap - 6.108);
                    double condht = vap > 6.108 ? d * factr : 0.0;
                    d = reltmp * (1.0 - 0.3 * melev / 10000.0) * factr;
                    double convht = airtmp > 32.0 ? d : 0.0;
                    d = sqrt(dull / 24.0);
                    double summer = py_max(0.45, 0.80 - 0.10 * d);
                    double winter = py_max(0.60, 0.85 - 0.07 * d);
                    albedo = ((int64_t)in->SEASONS[t]) ? summer : winter;
                    double kk = (1.0 - shade) * delt60;
                    double long1 = (shade * 0.26 * reltmp) + kk * (0.20 * reltmp - 6.6);
                    double long2 = (shade * 0.20 * reltmp) + kk * (0.17 * reltmp - 6.6);
                    double lng = reltmp > 0.0 ? long1 : long2;
                    if (lng < 0.0) lng *= skyclr;
                    double shrt = solrad * (1.0 - albedo) * (1.0 - shade);
                    mostht = (shrt + lng) / 203.2 + convht + condht;
                } else /* --- DEGDAY (SNOW.py:380) --- */
                    mostht = (in->KMELT[t] * delt / 1440.0) * (airtmp - tbase);
            }
            double rnsht = rainf > 0.0 ? reltmp * rainf / 144.0 : 0.0;
            double sumht = mostht + rnsht;
            if (snocov < 1.0) sumht = sumht * snocov;
            paktmp = neghts <= 0.0 ? 32.0 : 32.0 - neghts / (0.00695 * packf);

            /* --- COOLER (SNOW.py:391-399) --- */
            if (iregfg) {
                mneghs = reltmp > 0.0 ? 0.0 : -reltmp * 0.00695 * packf / 2.0;
                neght = paktmp <= airtmp ? 0.0 : 0.0007 * (paktmp - airtmp) * delt60;
            }
            if (sumht < 0.0) {
                if (paktmp > in->AIRTMP[t]) neghts = py_min(mneghs, neghts + neght);
                sumht = 0.0;
            }

            /* --- WARMUP (SNOW.py:402-426) --- */
            double rnfrz;
            if (neghts > 0.0) {
                if (sumht > 0.0) {
                    if (sumht > neghts) {
                        sumht -= neghts;
                        neghts = 0.0;
                    } else {
                        neghts -= sumht;
                        sumht = 0.0;
                    }
                }
                if (prain > 0.0) {
                    if (prain > neghts) {
                        rnfrz = neghts;
                        packf += rnfrz;
                        neghts = 0.0;
                    } else {
                        rnfrz = prain;
                        neghts -= prain;
                        packf += prain;
                    }
                    if (packf > pdepth) pdepth = packf;
                } else
                    rnfrz = 0.0;
            } else
                rnfrz = 0.0;

            /* --- MELTER (SNOW.py:428-441) --- */
            if (sumht >= packf) {
                melt = packf;
                packf = 0.0;
                pdepth = 0.0;
                packi = 0.0;
            } else if (sumht > 0.0) {
                melt = sumht;
                pdepth *= (1.0 - melt / packf);
                packf -= melt;
                if (packi > packf) packi = packf;
            } else
                melt = 0.0;

            /* --- LIQUID (SNOW.py:443-458) --- */
            if (iregfg && packf > 0.0) {
                rdenpf = packf / pdepth;
                if (rdenpf <= 0.6)
                    packwc = mwater;
                else {
                    double df = 3.0 - 3.33 * rdenpf;
                    packwc = df >= 0.0 ? mwater * df : 0.0;
                }
            }
            double pwsupy = packw + melt + prain - rnfrz;
            double mpws = packwc * packf;
            if ((pwsupy - mpws) > (0.01 * delt60)) {
                wyield = pwsupy - mpws;
                packw = mpws;
            } else {
                packw = pwsupy;
                wyield = 0.0;
            }

            /* --- ICING (SNOW.py:461-485) --- */
            if (icefg) {
                if (hr6ind < 1) {
                    if (hr6fg != 0) {
                        if (snocov < 1.0) {
                            double xlnem = -reltmp * 0.01;
                            if (xlnem > xlnmlt) xlnmlt = xlnem;
                        }
                        hr6fg = 0;
                    }
                } else
                    hr6fg = 1;
                if (wyield > 0.0 && xlnmlt > 0.0) {
                    double freeze;
                    if (wyield < xlnmlt) {
                        freeze = wyield;
                        xlnmlt -= wyield;
                        wyield = 0.0;
                    } else {
                        freeze = xlnmlt;
                        wyield -= xlnmlt;
                        xlnmlt = 0.0;
                    }
                    packf += freeze;
                    packi += freeze;
                    pdepth += freeze;
                }
            }

            /* --- GMELT (SNOW.py:487-508) --- */
            if (iregfg) {
                if (paktmp >= 32.0)
                    gmeltr = mgmelt;
                else if (paktmp > 5.0)
                    gmeltr = mgmelt * (1.0 - 0.03 * (32.0 - paktmp));
                else
                    gmeltr = mgmelt * 0.19;
            }
            if (packf <= gmeltr) {
                wyield = wyield + packf + packw;
                packf = packi = packw = 0.0;
                pdepth = 0.0;
                neghts = 0.0;
            } else {
                double d = 1.0 - (gmeltr / packf);
                packw += gmeltr;
                pdepth *= d;
                neghts *= d;
                packf -= gmeltr;
                packi = packi > gmeltr ? packi - gmeltr : 0.0;
            }

            /* --- pack bookkeeping / NOPACK (SNOW.py:511-537) --- */
            if (packf > 0.005) {
                rdenpf = packf / pdepth;
                paktmp = neghts == 0.0 ? 32.0 : 32.0 - neghts / (0.00695 * packf);
                snocov = packf < covinx ? packf / covinx : 1.0;
            } else {
                melt += packf;
                wyield += packf + packw;
                hr6fg = 1;
                packf = packi = packw = 0.0;
                pdepth = 0.0;
                rdenpf = NAN;
                covinx = 0.1 * covind;
                snocov = 0.0;
                xlnmlt = 0.0;
                mneghs = NAN;
                paktmp = 32.0;
                neghts = 0.0;
                dull = -1.0E30;
                albedo = -1.0E30;
                snowep = -1.0E30;
                vap = -1.0E30;
            }
        } else { /* SNOW.py:538-543 */
            prain = 0.0;
            snowe = 0.0;
            wyield = 0.0;
            melt = 0.0;
        }

        out->ALBEDO[t] = albedo;
        out->COVINX[t] = covinx;
        out->DEWTMP[t] = dewtmp;
        out->DULL[t] = dull;
        out->MELT[t] = melt;
        out->NEGHTS[t] = neghts;
        out->PACKF[t] = packf;
        out->PACKI[t] = packi;
        out->PACKW[t] = packw;
        out->PACK[t] = packf + packw;
        out->PAKTMP[t] = paktmp;
        out->PDEPTH[t] = pdepth;
        out->PRAIN[t] = prain;
        out->RAINF[t] = rainf;
        out->RDENPF[t] = rdenpf;
        out->SKYCLR[t] = skyclr;
        out->SNOCOV[t] = snocov;
        out->SNOTMP[t] = snotmp;
        out->SNOWE[t] = snowe;
        out->SNOWF[t] = snowf;
        out->WYIELD[t] = wyield;
        out->XLNMLT[t] = xlnmlt;
        if (metric) { /* SNOW.py:569-585: the stored series only; the state stays in English units */
            out->PAKTMP[t] = (paktmp * 0.555) - 17.777;
            out->DEWTMP[t] = (dewtmp * 0.555) - 17.777;
            out->SNOTMP[t] = (snotmp * 0.555) - 17.777;
            out->PACKF[t] = packf * 25.4;
            out->PACKI[t] = packi * 25.4;
            out->PACKW[t] = packw * 25.4;
            out->PACK[t] = out->PACKF[t] + out->PACKW[t];
            out->PDEPTH[t] = pdepth * 25.4;
            out->NEGHTS[t] = neghts * 25.4;
            out->COVINX[t] = covinx * 25.4;
            out->SNOWE[t] = snowe * 25.4;
            out->SNOWF[t] = snowf * 25.4;
            out->WYIELD[t] = wyield * 25.4;
            out->XLNMLT[t] = xlnmlt * 25.4;
            out->MELT[t] = melt * 25.4;
            out->PRAIN[t] = prain * 25.4;
        }
    }
    (void)satvap;
}

/* ============================================================================================== */
/* PWATER  (PWATER.py:102-769)                                                                     */
/* ============================================================================================== */
static const double UZRA[10] = {0.0, 1.25, 1.50, 1.75, 2.00, 2.10, 2.20, 2.25, 2.5, 4.0};
static const double INTGRL[10] = {0.0, 1.29, 1.58, 1.92, 2.36, 2.81, 3.41, 3.8, 7.1, 3478.};

/* numpy argmax(x < table) - 1  (PWATER.py:388,395): first k with x < table[k], minus one; none => -1 */
static int first_above_minus1(double x, const double *tab) {
    for (int k = 0; k < 10; ++k)
        if (x < tab[k]) return k - 1;
    return -1;
}

/* PROUTE (PWATER.py:696-769) */
static void proute(double psur, int rtopfg, double delt60, double dec, double src, double *surs_io,
                   double *suro_out, int64_t *errors) {
    double surs = *surs_io, suro;
    if (psur > 0.0002) {
        if (rtopfg != 1) {
            double ssupr = (psur - surs) / delt60;
            double surse = ssupr > 0.0 ? dec * pow(ssupr, 0.6) : 0.0;
            double sursnw = psur;
            suro = 0.0;
            int converged = 0;
            for (int count = 0; count < 100; ++count) {
                double ratio, fact;
                if (ssupr > 0.0) {
                    ratio = sursnw / surse;
                    fact = ratio <= 1.0 ? 1.0 + 0.6 * (ratio * (ratio * ratio)) : 1.6;
                } else {
                    ratio = 1.0e30;
                    fact = 1.6;
                }
                double ffact = (delt60 * src * pow(fact, 1.667)) * pow(sursnw, 1.667);
                double fsuro = ffact - suro;
                double dfact = -1.667 * ffact;
                double dfsuro = dfact / sursnw - 1.0;
                if (ratio <= 1.0) {
                    double dterm = dfact / (fact * surse) * 1.8 * (ratio * ratio);
                    dfsuro = dfsuro + dterm;
                }
                double dsuro = fsuro / dfsuro;
                suro = suro - dsuro;
                if (suro <= 1.0e-10) suro = 0.0;
                sursnw = psur - suro;
                double change = 0.0;
                if (fabs(suro) > 0.0) change = fabs(dsuro / suro);
                if (change < 0.01) {
                    converged = 1;
                    break;
                }
            }
            if (!converged) errors[6] += 1;
            surs = sursnw;
        } else {
            double ssupr = psur - surs;
            double sursm = (surs + psur) * 0.5;
            double dummy;
            if (ssupr > 0.0) {
                dummy = dec * pow(ssupr, 0.6);
                if (dummy > sursm) {
                    double surse = dummy;
                    double r = sursm / surse;
                    dummy = sursm * (1.0 + 0.6 * (r * (r * r)));
                } else
                    dummy = sursm * 1.6;
            } else
                dummy = sursm * 1.6;
            double tsuro = delt60 * src * pow(dummy, 1.667);
            suro = tsuro > psur ? psur : tsuro;
            surs = tsuro > psur ? 0.0 : psur - suro;
        }
    } else {
        suro = psur;
        surs = 0.0;
    }
    if (suro <= 1.0e-10) suro = 0.0;
    *surs_io = surs;
    *suro_out = suro;
}

void hsp2o_pwater(const pwater_ui *ui, const pwater_in *in, const pwater_out *out, int64_t nsteps,
                  int64_t *errors) {
    for (int i = 0; i < HSP2O_NERR_PWATER; ++i) errors[i] = 0;
    const double delt60 = ui->delt / 60.0;
    const int ICEFG = (int)ui->ICEFG, CSNOFG = (int)ui->CSNOFG, IFFCFG = (int)ui->IFFCFG,
              IFRDFG = (int)ui->IFRDFG, RTOPFG = (int)ui->RTOPFG, UZFG = (int)ui->UZFG, VLEFG = (int)ui->VLEFG;
    const double agwetp = ui->AGWETP, basetp = ui->BASETP, infexp = ui->INFEXP, infild = ui->INFILD,
                 lsur = ui->LSUR, slsur = ui->SLSUR, forest = ui->FOREST;
    /* PWATER.py:203-205 reads FZG/FZGL only when ICEFG; with ICEFG=0 and IFFCFG=2 the reference then uses an
       unassigned `fzgl` (PWATER.py:303), which Numba 0.65 materialises as 0.0 (observed; pinned by the golden
       case p_nosnow_iffc2).  Replicated, not fixed. */
    const double fzg = ICEFG ? ui->FZG : 0.0, fzgl = ICEFG ? ui->FZGL : 0.0;
    double agws = ui->AGWS, ceps = ui->CEPS, gwvs = ui->GWVS, ifws = ui->IFWS, lzs = ui->LZS, surs = ui->SURS,
           uzs = ui->UZS;

    double rlzrat = -1.0E30, lzfrac = -1.0E30, rparm = -1.0E30;
    if (agws < 0.0) agws = 0.0;
    double msupy = 0.0, dec = NAN, src = NAN, kifw = NAN, ifwk2 = NAN, ifwk1 = NAN;
    double petadj = 0.0; /* undefined in the reference until the first HRFG step, which is step 0 */

    for (int64_t t = 0; t < nsteps; ++t) {
        const double oldmsupy = msupy;
        const int dayfg = (int)(int64_t)in->DAYFG[t];
        const int hrfg = (int)(int64_t)in->HRFG[t];
        double inffac = 1.0;
        const double kgw = 1.0 - pow(in->AGWRC[t], delt60 / 24.0); /* PWATER.py:247 */
        const double lzetp = in->LZETP[t], cepsc = in->CEPSC[t], uzsn = in->UZSN[t];
        /* PWATER.py:215; a metric run multiplies the product by 0.0394 (:239): INFILT_MFAC, 1.0 otherwise (x * 1.0 == x) */
        const double infilt = (in->INFILT[t] * delt60) * ui->INFILT_MFAC;
        const double kvary = in->KVARY[t], lzsn = in->LZSN[t];
        const double petinp = in->PETINP[t];
        double supy, pet, petadj_out = 0.0;

        /* supply / PET (PWATER.py:279-303) */
        if (CSNOFG) {
            double airtmp = in->AIRTMP[t], petmax = in->PETMAX[t], petmin = in->PETMIN[t];
            double snocov = in->SNOCOV[t];
            supy = in->RAINF[t] * (1.0 - snocov) + in->WYIELD[t];
            if (hrfg) {
                petadj = (1.0 - forest) * (1.0 - snocov) + forest;
                if ((airtmp < petmax) && (petadj > 0.5)) petadj = 0.5;
                if (airtmp < petmin) petadj = 0.0;
                petadj_out = petadj;
            }
            pet = petinp * petadj;
            if (ICEFG) inffac = py_max(fzgl, 1.0 - fzg * in->PACKI[t]);
        } else {
            supy = in->PREC[t];
            pet = petinp;
        }
        if (IFFCFG == 2) inffac = in->LGTMP[t] <= 0.0 ? fzgl : 1.0;

        /* ICEPT (PWATER.py:305-312) */
        ceps = ceps + supy + 0.0;
        double cepo = 0.0;
        if (ceps > cepsc) {
            cepo = ceps - cepsc;
            ceps = cepsc;
        }
        const double suri = cepo + in->SURLI[t];
        msupy = suri + surs + 0.0;
        double lzrat = lzs / lzsn;
        double suro, ifwi, infil, uzi;

        if (msupy <= 0.0) {
            surs = 0.0;
            suro = 0.0;
            ifwi = 0.0;
            infil = 0.0;
            uzi = 0.0;
        } else { /* SURFAC / DISPOS (PWATER.py:326-437) */
            double ibar = infilt / pow(lzrat, infexp);
            if (inffac < 1.0) ibar = ibar * inffac;
            double imax = ibar * infild;
            double imin = ibar - (imax - ibar);
            if (dayfg || oldmsupy == 0.0) {
                double dummy = in->NSUR[t] * lsur;
                dec = 0.00982 * pow(dummy / sqrt(slsur), 0.6);
                src = 1020.0 * (sqrt(slsur) / dummy);
            }
            double ratio = py_max(1.0001, in->INTFW[t] * exp2(lzrat));
            double under, over;
            if (msupy <= imin) {
                under = msupy;
                over = 0.0;
            } else if (msupy > imax) {
                under = (imin + imax) * 0.5;
                over = msupy - under;
            } else {
                over = ((msupy - imin) * (msupy - imin)) * 0.5 / (imax - imin);
                under = msupy - over;
            }
            infil = under;
            if (over <= 0.0) {
                surs = 0.0;
                suro = 0.0;
                ifwi = 0.0;
                uzi = 0.0;
            } else {
                double pdro = over, uzfrac;
                if (UZFG) { /* UZINF2 (PWATER.py:366-379) */
                    double uzrat = uzs / uzsn;
                    if (uzrat < 2.0) {
                        double k1 = 3.0 - uzrat;
                        uzfrac = 1.0 - (uzrat * 0.5) * pow(1.0 / (1.0 + k1), k1);
                    } else {
                        double k2 = (2.0 * uzrat) - 3.0;
                        uzfrac = pow(1.0 / (1.0 + k2), k2);
                    }
                    uzi = pdro * uzfrac;
                } else { /* UZINF table look-up (PWATER.py:381-404) */
                    double uzraa = uzs / uzsn;
                    int kk = first_above_minus1(uzraa, UZRA);
                    if (kk == -1) {
                        kk = 8;
                        errors[1] += 1;
                    }
                    double intga = INTGRL[kk] + (INTGRL[kk + 1] - INTGRL[kk]) * (uzraa - UZRA[kk]) / (UZRA[kk + 1] - UZRA[kk]);
                    double intgb = (pdro / uzsn) + intga;
                    kk = first_above_minus1(intgb, INTGRL);
                    if (kk == -1) {
                        errors[2] += 1;
                        kk = 8;
                    }
                    double uzrab = UZRA[kk] + (UZRA[kk + 1] - UZRA[kk]) * (intgb - INTGRL[kk]) / (INTGRL[kk + 1] - INTGRL[kk]);
                    uzi = (uzrab - uzraa) * uzsn;
                    if (uzi < -1.0e-3) errors[7] += 1;
                    uzi = py_max(0.0, uzi);
                }
                if (uzi > pdro) uzi = pdro;
                uzfrac = uzi / pdro;
                double iimin = imin * ratio, iimax = imax * ratio, over2;
                if (msupy <= iimin)
                    over2 = 0.0;
                else if (msupy > iimax)
                    over2 = msupy - (iimin + iimax) * 0.5;
                else
                    over2 = ((msupy - iimin) * (msupy - iimin)) * 0.5 / (iimax - iimin);
                double psur = over2;
                double pifwi = pdro - psur;
                ifwi = pifwi * (1.0 - uzfrac);
                if (psur <= 0.0) {
                    surs = 0.0;
                    suro = 0.0;
                } else {
                    psur = psur * (1.0 - uzfrac);
                    proute(psur, RTOPFG, delt60, dec, src, &surs, &suro, errors);
                }
            }
        }

        /* INTFLW (PWATER.py:440-455) */
        if (dayfg) {
            kifw = -log(in->IRC[t]) / (24.0 / delt60);
            ifwk2 = 1.0 - exp(-kifw);
            ifwk1 = 1.0 - (ifwk2 / kifw);
        }
        double inflo = ifwi + in->IFWLI[t];
        double value = inflo + ifws;
        double ifwo;
        if (value > 0.00002) {
            ifwo = (ifwk1 * inflo) + (ifwk2 * ifws);
            ifws = value - ifwo;
        } else {
            ifwo = 0.0;
            ifws = 0.0;
            uzs = uzs + value;
        }

        /* UZONE (PWATER.py:457-468) */
        double uzrat = uzs / uzsn;
        uzs = uzs + uzi + in->UZLI[t] + 0.0;
        double perc = 0.0;
        if (uzrat - lzrat > 0.01) {
            double d = uzrat - lzrat;
            perc = 0.1 * infilt * inffac * uzsn * (d * (d * d));
            if (perc > uzs) {
                perc = uzs;
                uzs = 0.0;
            } else
                uzs -= perc;
        }
        double iperc = perc + infil + in->LZLI[t];

        /* LZONE (PWATER.py:473-487) */
        double lperc = iperc + 0.0;
        double lzi = 0.0;
        if (lperc > 0.0) {
            if (fabs(lzrat - rlzrat) > 0.02 || IFRDFG) {
                rlzrat = lzrat;
                if (lzrat <= 1.0) {
                    double indx = 2.5 - 1.5 * lzrat;
                    lzfrac = IFRDFG ? 1.0 : 1.0 - lzrat * pow(1.0 / (1.0 + indx), indx);
                } else {
                    double indx = 1.5 * lzrat - 0.5;
                    double exfact = -1.0 * IFRDFG;
                    lzfrac = IFRDFG ? exp(exfact * (lzrat - 1.0)) : pow(1.0 / (1.0 + indx), indx);
                }
            }
            lzi = lzfrac * lperc;
            lzs += lzi;
        }
        double gwi = iperc + 0.0 - lzi;

        /* GWATER (PWATER.py:492-528) */
        double igwi = 0.0, agwi = 0.0;
        if (gwi > 0.0) {
            igwi = in->DEEPFR[t] * gwi;
            agwi = gwi - igwi;
        }
        double ainflo = agwi + in->AGWLI[t] + 0.0;
        double agwo = 0.0;
        if (kvary > 0.0) {
            gwvs += ainflo;
            if (dayfg) gwvs = gwvs > 0.0001 ? gwvs * 0.97 : 0.0;
            if (agws > 1.0e-20) {
                agwo = kgw * (1.0 + kvary * gwvs) * agws;
                double avail = ainflo + agws;
                if (agwo > avail) {
                    errors[3] += 1;
                    agwo = avail;
                }
            }
        } else if (agws > 1.0e-20)
            agwo = kgw * agws;
        if (agwo < 1.0e-12) agwo = 0.0;
        agws = agws + (ainflo - agwo);
        if (agws < 0.0) {
            errors[8] += 1;
            agws = 0.0;
        }

        /* EVAPT (PWATER.py:539-626) */
        double rempet = pet, taet = 0.0, baset = 0.0;
        if (rempet > 0.0 && basetp > 0.0) {
            double baspet = basetp * rempet;
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
        if (rempet > 0.0 && ceps > 0.0) {
            if (rempet > ceps) {
                cepe = ceps;
                ceps = 0.0;
            } else {
                cepe = rempet;
                ceps -= cepe;
            }
            taet += cepe;
            rempet -= cepe;
        }
        double uzet = 0.0;
        if (rempet > 0.0) {
            if (uzs > 0.001) {
                uzrat = uzs / uzsn;
                double uzpet = uzrat > 2.0 ? rempet : 0.5 * uzrat * rempet;
                if (uzpet > uzs) {
                    uzet = uzs;
                    uzs = 0.0;
                } else {
                    uzet = uzpet;
                    uzs -= uzet;
                }
            }
            taet += uzet;
            rempet -= uzet;
        }
        double agwet = 0.0;
        if (rempet > 0.0 && agwetp > 0.0) {
            double gwpet = rempet * agwetp;
            if (gwpet > agws) {
                agwet = agws;
                agws = 0.0;
            } else {
                agwet = gwpet;
                agws -= agwet;
            }
            if (fabs(kvary) > 0.0) gwvs -= agwet;
            taet += agwet;
            rempet -= agwet;
        }
        if (dayfg) {
            lzrat = lzs / lzsn;
            rparm = lzetp <= 0.99999 ? 0.25 / (1.0 - lzetp) * lzrat * delt60 / 24.0 : 1.0e10;
        }
        double lzet = 0.0;
        if (rempet > 0.0 && lzs > 0.02) {
            double lzpet;
            if (lzetp >= 0.99999)
                lzpet = rempet * lzetp;
            else if (VLEFG <= 1) {
                lzpet = rempet > rparm ? 0.5 * rparm : rempet * (1.0 - rempet / (2.0 * rparm));
                if (lzetp < 0.5) lzpet = lzpet * 2.0 * lzetp;
            } else
                lzpet = lzrat < 1.0 ? lzetp * lzrat * rempet : lzetp * rempet;
            lzet = lzpet < (lzs - 0.02) ? lzpet : lzs - 0.02;
            lzs -= lzet;
            taet += lzet;
            rempet -= lzet;
        }
        const double tgws = agws;

        out->AGWET[t] = agwet;
        out->AGWI[t] = agwi;
        out->AGWO[t] = agwo;
        out->AGWS[t] = agws;
        out->BASET[t] = baset;
        out->CEPE[t] = cepe;
        out->CEPS[t] = ceps;
        out->GWVS[t] = gwvs;
        out->IFWI[t] = ifwi;
        out->IFWO[t] = ifwo;
        out->IFWS[t] = ifws;
        out->IGWI[t] = igwi;
        out->INFFAC[t] = inffac;
        out->INFIL[t] = infil;
        out->LZET[t] = lzet;
        out->LZI[t] = lzi;
        out->LZS[t] = lzs;
        out->PERC[t] = perc;
        out->PERO[t] = suro + ifwo + agwo;
        out->PERS[t] = ceps + surs + ifws + uzs + lzs + tgws;
        out->PET[t] = pet;
        out->PETADJ[t] = petadj_out;
        out->SUPY[t] = supy;
        out->SURI[t] = suri;
        out->SURO[t] = suro;
        out->SURS[t] = surs;
        out->TAET[t] = taet;
        out->TGWS[t] = tgws;
        out->UZET[t] = uzet;
        out->UZI[t] = uzi;
        out->UZS[t] = uzs;
    }
}

/* ============================================================================================== */
/* IWATER  (IWATER.py:75-284)                                                                      */
/* ============================================================================================== */
void hsp2o_iwater(const iwater_ui *ui, const iwater_in *in, const iwater_out *out, int64_t nsteps,
                  int64_t *errors) {
    errors[0] = 0;
    const double delt60 = ui->delt / 60.0;
    const double lsur = ui->LSUR, slsur = ui->SLSUR;
    const int RTLIFG = (int)ui->RTLIFG, CSNOFG = (int)ui->CSNOFG, RTOPFG = (int)ui->RTOPFG;
    double rets = ui->RETS, surs = ui->SURS;
    double msupy = surs;
    double dec = NAN, src = NAN, surse = NAN, ssupr = NAN, dummy = NAN;
    double petadj = 0.0;

    for (int64_t t = 0; t < nsteps; ++t) {
        const double oldmsupy = msupy;
        const double petinp = in->PETINP[t];
        double supy, pet, petadj_out = 0.0;
        if (CSNOFG) { /* IWATER.py:152-167 */
            double airtmp = in->AIRTMP[t], petmax = in->PETMAX[t], petmin = in->PETMIN[t];
            double snocov = in->SNOCOV[t];
            supy = in->RAINF[t] * (1.0 - snocov) + in->WYIELD[t];
            if ((int64_t)in->HRFG[t]) {
                petadj = 1.0 - snocov;
                if (airtmp < petmax) {
                    if (airtmp < petmin) petadj = 0.0;
                    if (petadj > 0.5) petadj = 0.5;
                }
                petadj_out = petadj;
            }
            pet = petinp * petadj;
        } else {
            supy = in->PREC[t];
            pet = petinp;
        }
        const double surli = in->SURLI[t];
        const double retsc = in->RETSC[t];
        double suri, reto;
        /* RETN (IWATER.py:175-196) */
        double reti = RTLIFG ? supy + surli : supy;
        rets += reti;
        if (rets > retsc) {
            reto = rets - retsc;
            rets = retsc;
        } else
            reto = 0.0;
        suri = RTLIFG ? reto : reto + surli;
        msupy = suri + surs;

        double suro = 0.0;
        const int hr1 = (int)(int64_t)in->HR1FG[t];
        if (msupy > 0.0002) {
            if (oldmsupy == 0.0 || hr1) { /* IWATER.py:204-207, 221-224 */
                dummy = in->NSUR[t] * lsur;
                dec = 0.00982 * pow(dummy / sqrt(slsur), 0.6);
                src = 1020.0 * sqrt(slsur) / dummy;
            }
            if (RTOPFG) { /* ARM/NPS closed form (IWATER.py:202-218); note exponent 1.67 */
                double sursm = (surs + msupy) * 0.5;
                dummy = sursm * 1.6;
                if (suri > 0.0) {
                    double d = dec * pow(suri, 0.6);
                    if (d > sursm) {
                        surse = d;
                        double r = sursm / surse;
                        dummy = sursm * (1.0 + 0.6 * (r * (r * r)));
                    }
                }
                double tsuro = delt60 * src * pow(dummy, 1.67);
                suro = tsuro > msupy ? msupy : tsuro;
                surs = tsuro > msupy ? 0.0 : msupy - suro;
            } else { /* Newton (IWATER.py:219-254); no zero guards, TOLERANCE 0.01 */
                ssupr = suri / delt60;
                surse = ssupr > 0.0 ? dec * pow(ssupr, 0.6) : 0.0;
                double sursnw = msupy;
                suro = 0.0;
                int converged = 0;
                for (int count = 0; count < 100; ++count) {
                    double ratio, fact;
                    if (ssupr > 0.0) {
                        ratio = sursnw / surse;
                        fact = ratio <= 1.0 ? 1.0 + 0.6 * (ratio * (ratio * ratio)) : 1.6;
                    } else {
                        fact = 1.6;
                        ratio = 1e30;
                    }
                    double ffact = (delt60 * src * pow(fact, 1.667)) * pow(sursnw, 1.667);
                    double fsuro = ffact - suro;
                    double dfact = -1.667 * ffact;
                    double dfsuro = dfact / sursnw - 1.0;
                    if (ratio <= 1.0) dfsuro += (dfact / (fact * surse)) * 1.8 * (ratio * ratio);
                    double dsuro = fsuro / dfsuro;
                    suro = suro - dsuro;
                    sursnw = msupy - suro;
                    if (fabs(dsuro / suro) < 0.01) {
                        converged = 1;
                        break;
                    }
                }
                if (!converged) errors[0] = errors[0] + 1;
                surs = sursnw;
            }
        } else {
            suro = msupy;
            surs = 0.0;
        }

        /* EVRETN (IWATER.py:259-269) */
        double impev;
        if (rets > 0.0) {
            if (pet > rets) {
                impev = rets;
                rets = 0.0;
            } else {
                impev = pet;
                rets -= impev;
            }
        } else
            impev = 0.0;

        out->IMPEV[t] = impev;
        out->PET[t] = pet;
        out->PETADJ[t] = petadj_out;
        out->RETS[t] = rets;
        out->SUPY[t] = supy;
        out->SURI[t] = suri;
        out->SURO[t] = suro;
        out->SURS[t] = surs;
    }
    (void)ssupr;
}

/* ============================================================================================== */
/* SEDMNT  (SEDMNT.py:52-264)                                                                      */
/* ============================================================================================== */
void hsp2o_sedmnt(const sedmnt_ui *ui, const sedmnt_in *in, const sedmnt_out *out, int64_t nsteps) {
    const double delt = ui->delt, delt60 = delt / 60.0;
    const double VSIVFG = ui->VSIVFG; /* compared as a float in the reference */
    const int SDOPFG = ui->SDOPFG != 0.0, CSNOFG = (int)ui->CSNOFG;
    const double smpf = ui->SMPF, krer = ui->KRER, jrer = ui->JRER, affix = ui->AFFIX, kser = ui->KSER,
                 jser = ui->JSER, kger = ui->KGER, jger = ui->JGER;
    const int metric = ui->uunits == 2.0; /* SEDMNT.py:60,113-117: NVSI series, SURO / SURS (the wrapper) and DETS on entry */
    double nvsi = ui->NVSI * delt / 1440.;
    double cover = in->COVERI[0];
    double dets = ui->DETS;
    if (metric) dets = dets * 1.10231 / 2.471;
    int drydfg = 1;
    double wssd = 0.0, scrsd = 0.0, sosed = 0.0;

    for (int64_t t = 0; t < nsteps; ++t) {
        const int dayfg = (int)(int64_t)in->DAYFG[t];
        const double rain = in->RAIN[t], prec = in->PREC[t], suro = in->SURO[t], surs = in->SURS[t];
        const double snocov = in->SNOCOV[t], slsed = in->SLSED[t];
        if (dayfg) cover = in->COVERI[t];

        /* DETACH (SEDMNT.py:161-172) */
        if (rain > 0.0) {
            double cr;
            if (CSNOFG == 1)
                cr = snocov > 0.0 ? cover + (1.0 - cover) * snocov : cover;
            else
                cr = cover;
            double det = delt60 * (1.0 - cr) * smpf * krer * pow(rain / delt60, jrer);
            dets = dets + det;
        }
        if (dayfg) { /* SEDMNT.py:175-179 */
            nvsi = in->NVSI[t] * delt / 1440.;
            if (metric) nvsi = nvsi * 2.205 / 2.471; /* SEDMNT.py:79,114: NVSI = ts['NVSI'] * delt / 1440., then * 2.205 / 2.471 */
            if (VSIVFG == 2 && drydfg) {
                double d = nvsi * (1440. / delt);
                dets = dets + d;
            }
        }
        double d = slsed;
        if (VSIVFG <= 2) d = d + nvsi;
        dets = dets + d;

        if (SDOPFG) { /* SOSED method 1 (SEDMNT.py:188-212) */
            if (suro > 0.0) {
                double arg = surs + suro;
                double stcap = delt60 * kser * pow(arg / delt60, jser);
                if (stcap > dets)
                    wssd = dets * suro / arg;
                else
                    wssd = stcap * suro / arg;
                dets = dets - wssd;
                scrsd = delt60 * kger * pow(arg / delt60, jger);
                scrsd = scrsd * suro / arg;
                sosed = wssd + scrsd;
            } else {
                wssd = 0.0;
                scrsd = 0.0;
                sosed = 0.0;
            }
        } else { /* SOSED method 2 (SEDMNT.py:214-233) */
            if (suro > 0.0) {
                double stcap = delt60 * kser * pow(suro / delt60, jser);
                if (stcap > dets) {
                    wssd = dets;
                    dets = 0.0;
                } else {
                    wssd = stcap;
                    dets = dets - wssd;
                }
                scrsd = delt60 * kger * pow(suro / delt60, jger);
                sosed = wssd + scrsd;
            } else {
                wssd = 0.0;
                scrsd = 0.0;
                sosed = 0.0;
            }
        }
        if (dayfg) { /* ATACH (SEDMNT.py:237-248) */
            if (drydfg) dets = dets * (1.0 - affix);
            drydfg = 1;
        }
        if (prec > 0) drydfg = 0;

        out->DETS[t] = dets;
        out->WSSD[t] = wssd;
        out->SCRSD[t] = scrsd;
        out->SOSED[t] = sosed;
        out->COVER[t] = cover;
        if (metric) { /* SEDMNT.py:258-262: the stored series only */
            out->DETS[t] = dets / 1.10231 * 2.471;
            out->WSSD[t] = wssd / 1.10231 * 2.471;
            out->SCRSD[t] = scrsd / 1.10231 * 2.471;
            out->SOSED[t] = sosed / 1.10231 * 2.471;
        }
    }
}

/* ============================================================================================== */
/* PSTEMP  (PSTEMP.py:62-217); metric (uunits == 2): states and tables are deg C already, AIRTMP is  */
/* taken to Celsius as an array on entry (:110-111), nothing is converted back on exit (:205-215)     */
/* ============================================================================================== */
void hsp2o_pstemp(const pstemp_ui *ui, const pstemp_in *in, const pstemp_out *out, int64_t nsteps,
                  int64_t *errors) {
    for (int i = 0; i < HSP2O_NERR_PSTEMP; ++i) errors[i] = 0;
    const double TSOPFG = ui->TSOPFG; /* float compare in the reference */
    const int AIRTFG = (int)ui->AIRTFG;
    const int metric = ui->uunits == 2.0; /* PSTEMP.py:70 */
    double airtc = metric ? ui->AIRTC : (ui->AIRTC - 32.0) * 0.555; /* PSTEMP.py:97-101 */
    double sltmp = metric ? ui->SLTMP : (ui->SLTMP - 32.0) * 0.555;
    double ultmp = metric ? ui->ULTMP : (ui->ULTMP - 32.0) * 0.555;
    double lgtmp = metric ? ui->LGTMP : (ui->LGTMP - 32.0) * 0.555;
    /* metric: AIRTMP = (AIRTMP - 32.0) * 0.555 as an array (:110-111), airts = AIRTMP[0] of THAT array (:128) */
    double airts = metric ? (in->AIRTMP[0] - 32.0) * 0.555 : in->AIRTMP[0];
    (void)airtc;

    for (int64_t t = 0; t < nsteps; ++t) {
        const int hrfg = (int)(int64_t)in->HRFG[t];
        const double airtmp = metric ? (in->AIRTMP[t] - 32.0) * 0.555 : in->AIRTMP[t];
        const double airtcs = metric ? airts : (airts - 32.0) * 0.555; /* PSTEMP.py:136-142 */
        airtc = metric ? airtmp : (airtmp - 32.0) * 0.555;
        if (hrfg) {
            double aslt = metric ? in->ASLT[t] : (in->ASLT[t] - 32.0) * 0.555; /* PSTEMP.py:146-149 */
            double bslt = in->BSLT[t];
            sltmp = aslt + bslt * airtc;
        }
        if (TSOPFG == 1) {
            if (hrfg) {
                double ault = metric ? in->ULTP1[t] : (in->ULTP1[t] - 32.0) * 0.5555; /* PSTEMP.py:120-122 */
                double bult = in->ULTP2[t];
                ultmp = ault + bult * airtc;
                lgtmp = metric ? in->LGTP1[t] : (in->LGTP1[t] - 32.0) * 0.5555; /* PSTEMP.py:123 */
            }
        } else {
            double ulsmo = in->ULTP1[t];
            double ultdif = metric ? in->ULTP2[t] : 0.555 * in->ULTP2[t]; /* PSTEMP.py:125 */
            double ultmps = ultmp;
            ultmp = ultmps + ulsmo * (airtcs + ultdif - ultmps);
            double lgsmo = in->LGTP1[t];
            double lgtdif = metric ? in->LGTP2[t] : 0.555 * in->LGTP2[t]; /* PSTEMP.py:126 */
            double lgtmps = lgtmp;
            if (TSOPFG == 0)
                lgtmp = lgtmps + lgsmo * (airtcs + lgtdif - lgtmps);
            else
                lgtmp = lgtmps + lgsmo * (ultmp + lgtdif - lgtmps);
        }
        if (sltmp < -100) { sltmp = -100; errors[0] += 1; }
        if (sltmp > 100) { sltmp = 100; errors[3] += 1; }
        if (ultmp < -100) { ultmp = -100; errors[1] += 1; }
        if (ultmp > 100) { ultmp = 100; errors[4] += 1; }
        if (lgtmp < -100) { lgtmp = -100; errors[2] += 1; }
        if (lgtmp > 100) { lgtmp = 100; errors[5] += 1; }
        if (AIRTFG == 0) airts = airtmp;

        out->AIRTC[t] = metric ? airtc : (airtc * 9.0 / 5.0) + 32.0; /* PSTEMP.py:205-215 */
        out->SLTMP[t] = metric ? sltmp : (sltmp * 9.0 / 5.0) + 32.0;
        out->ULTMP[t] = metric ? ultmp : (ultmp * 9.0 / 5.0) + 32.0;
        out->LGTMP[t] = metric ? lgtmp : (lgtmp * 9.0 / 5.0) + 32.0;
    }
}

/* ============================================================================================== */
/* PWTGAS  (PWTGAS.py:65-352); metric (uunits == 2): ELEV, the soil temperatures and the flows are    */
/* converted on entry by the wrapper (:90-95,167-174), the stored series here (:310-350)              */
/* ============================================================================================== */
void hsp2o_pwtgas(const pwtgas_ui *ui, const pwtgas_in *in, const pwtgas_out *out, int64_t nsteps) {
    const int CSNOFG = (int)ui->CSNOFG;
    const double slifac = ui->SLIFAC, ilifac = ui->ILIFAC, alifac = ui->ALIFAC;
    const double gdvfg = ui->GDVFG;
    const int metric = ui->uunits == 2.0;
    const double elevgc = pow((288.0 - 0.00198 * ui->ELEV) / 288.0, 5.256);
    const double EFACTA = 407960., MFACTA = 0.2266;

    for (int64_t t = 0; t < nsteps; ++t) {
        const int dayfg = (int)(int64_t)in->DAYFG[t];
        const double suro = in->SURO[t], wyield = in->WYIELD[t], ifwo = in->IFWO[t], agwo = in->AGWO[t];
        const double surli = in->SURLI[t], ifwli = in->IFWLI[t], agwli = in->AGWLI[t];
        const double sltmp = in->SLTMP[t], slitmp = in->SLITMP[t], slidox = in->SLIDOX[t], slico2 = in->SLICO2[t];
        const double ultmp = in->ULTMP[t];
        double idoxp = in->IDOXP[t], ico2p = in->ICO2P[t], adoxp = in->ADOXP[t], aco2p = in->ACO2P[t];

        double sotmp = -1.0e30, sodox = -1.0e30, soco2 = -1.0e30;
        if (suro > 0.0) {
            sotmp = (sltmp - 32.0) * 5. / 9.;
            if (sotmp < 0.5) sotmp = 0.5;
            if (CSNOFG) {
                if (wyield > 0.0) sotmp = 0.5;
            }
            double dummy = sotmp * (0.007991 - 0.77774E-4 * sotmp);
            sodox = (14.652 + sotmp * (-0.41022 + dummy)) * elevgc;
            double abstmp = sotmp + 273.16;
            dummy = 2385.73 / abstmp - 14.0184 + 0.0152642 * abstmp;
            soco2 = pow(10.0, dummy) * 3.16e-04 * elevgc * 12000.0;
            if (surli > 0.0 && slifac > 0.0) {
                if (slitmp >= -1.0e10) sotmp = slitmp * slifac + sotmp * (1.0 - slifac);
                if (slidox >= 0.0) sodox = slidox * slifac + sodox * (1.0 - slifac);
                if (slico2 >= 0.0) soco2 = slico2 * slifac + soco2 * (1.0 - slifac);
            }
        }
        double ilitmp = -1.0e30, ilidox = -1.0e30, ilico2 = -1.0e30;
        if (ifwli > 0.0) {
            ilitmp = in->ILITMP[t];
            ilidox = in->ILIDOX[t];
            ilico2 = in->ILICO2[t];
        }
        if (dayfg) {
            idoxp = in->IDOXP[t];
            ico2p = in->ICO2P[t];
        }
        double iotmp = -1.0e30, iodox = -1.0e30, ioco2 = -1.0e30;
        if (ifwo > 0.0) {
            iotmp = (ultmp - 32.0) * 5. / 9.;
            if (iotmp < 0.5) iotmp = 0.5;
            iodox = idoxp;
            ioco2 = ico2p;
            if (ifwli > 0.0 && ilifac > 0.0) {
                if (ilitmp >= -1.0e10) iotmp = ilitmp * ilifac + iotmp * (1.0 - ilifac);
                if (ilidox >= 0.0) iodox = ilidox * ilifac + iodox * (1.0 - ilifac);
                if (ilico2 >= 0.0) ioco2 = ilico2 * ilifac + ioco2 * (1.0 - ilifac);
            }
        }
        double alitmp = -1.0e30, alidox = -1.0e30, alico2 = -1.0e30;
        if (agwli > 0.0) {
            alitmp = in->ALITMP[t];
            alidox = in->ALIDOX[t];
            alico2 = in->ALICO2[t];
        }
        if (dayfg) {
            if (gdvfg != 0.0) {
                adoxp = in->ADOXP[t];
                aco2p = in->ACO2P[t];
            }
        }
        double aotmp = -1.0e30, aodox = -1.0e30, aoco2 = -1.0e30;
        if (agwo > 0.0) {
            aotmp = (in->LGTMP[t] - 32.0) * 5. / 9.;
            if (aotmp < 0.5) aotmp = 0.5;
            aodox = adoxp;
            aoco2 = aco2p;
            if (agwli > 0.0 && alifac > 0.0) {
                if (alitmp >= -1.0e10) aotmp = alitmp * alifac + aotmp * (1.0 - alifac);
                if (alidox >= 0.0) aodox = alidox * alifac + aodox * (1.0 - alifac);
                if (alico2 >= 0.0) aoco2 = alico2 * alifac + aoco2 * (1.0 - alifac);
            }
        }
        double soht = sotmp * suro * EFACTA, ioht = iotmp * ifwo * EFACTA, aoht = aotmp * agwo * EFACTA;
        out->POHT[t] = soht + ioht + aoht;
        double sodoxm = sodox * suro * MFACTA, iodoxm = iodox * ifwo * MFACTA, aodoxm = aodox * agwo * MFACTA;
        out->PODOXM[t] = sodoxm + iodoxm + aodoxm;
        double soco2m = soco2 * suro * MFACTA, ioco2m = ioco2 * ifwo * MFACTA, aoco2m = aoco2 * agwo * MFACTA;
        out->POCO2M[t] = soco2m + ioco2m + aoco2m;

        out->SOTMP[t] = (sotmp > -1e28 && !metric) ? (sotmp * 9.0 / 5.0) + 32.0 : sotmp;
        out->IOTMP[t] = (iotmp > -1e28 && !metric) ? (iotmp * 9.0 / 5.0) + 32.0 : iotmp;
        out->AOTMP[t] = (aotmp > -1e28 && !metric) ? (aotmp * 9.0 / 5.0) + 32.0 : aotmp;
        out->SODOX[t] = sodox;
        out->SOCO2[t] = soco2;
        out->IODOX[t] = iodox;
        out->IOCO2[t] = ioco2;
        out->AODOX[t] = aodox;
        out->AOCO2[t] = aoco2;
        out->SOHT[t] = soht;
        out->IOHT[t] = ioht;
        out->AOHT[t] = aoht;
        out->SODOXM[t] = sodoxm;
        out->SOCO2M[t] = soco2m;
        out->IODOXM[t] = iodoxm;
        out->IOCO2M[t] = ioco2m;
        out->AODOXM[t] = aodoxm;
        out->AOCO2M[t] = aoco2m;
        if (metric) { /* PWTGAS.py:338-350 */
            out->SOHT[t] = out->SOHT[t] / 3.96567 * 2.471;
            out->IOHT[t] = out->IOHT[t] / 3.96567 * 2.471;
            out->AOHT[t] = out->AOHT[t] / 3.96567 * 2.471;
            out->POHT[t] = out->POHT[t] / 3.96567 * 2.471;
            out->SODOXM[t] = out->SODOXM[t] / 2.205 * 2.471;
            out->SOCO2M[t] = out->SOCO2M[t] / 2.205 * 2.471;
            out->IODOXM[t] = out->IODOXM[t] / 2.205 * 2.471;
            out->IOCO2M[t] = out->IOCO2M[t] / 2.205 * 2.471;
            out->AODOXM[t] = out->AODOXM[t] / 2.205 * 2.471;
            out->AOCO2M[t] = out->AOCO2M[t] / 2.205 * 2.471;
            out->PODOXM[t] = out->PODOXM[t] / 2.205 * 2.471;
            out->POCO2M[t] = out->POCO2M[t] / 2.205 * 2.471;
        }
    }
}

/* ============================================================================================== */
/* SOLIDS  (SOLIDS.py:52-146), IMPLND.  SLSLD is read by the reference but unused.  Metric: every    */
/* conversion is elementwise on entry / exit (:84-88,143-145) and lives in oracle/wrappers.py          */
/* ============================================================================================== */
void hsp2o_solids(const solids_ui *ui, const solids_in *in, const solids_out *out, int64_t nsteps) {
    const double delt60 = ui->delt60, keim = ui->KEIM, jeim = ui->JEIM;
    const double SDOPFG = ui->SDOPFG; /* compared as a float in the reference (SOLIDS.py:103) */
    double slds = ui->SLDS;
    int drydfg = 1; /* SOLIDS.py:75 */
    for (int64_t t = 0; t < nsteps; ++t) {
        const double suro = in->SURO[t], surs = in->SURS[t], prec = in->PREC[t];
        const double accsdp = in->ACCSDP[t], remsdp = in->REMSDP[t];
        const int dayfg = (int)in->DAYFG[t];
        double sosld;
        if (SDOPFG == 1) { /* method 1 (SOLIDS.py:103-114) */
            if (suro > 0.0) {
                const double arg = surs + suro;
                const double stcap = delt60 * keim * pow(arg / delt60, jeim);
                if (stcap > slds)
                    sosld = slds * suro / arg;
                else
                    sosld = stcap * suro / arg;
                slds = slds - sosld;
            } else
                sosld = 0.0;
        } else { /* method 2 (SOLIDS.py:115-126) */
            if (suro > 0.0) {
                const double stcap = delt60 * keim * pow(suro / delt60, jeim);
                if (stcap > slds) {
                    sosld = slds;
                    slds = 0.0;
                } else {
                    sosld = stcap;
                    slds = slds - sosld;
                }
            } else
                sosld = 0.0;
        }
        if (dayfg == 0) { /* SOLIDS.py:129-142 */
            if (prec > 0.0) drydfg = 0;
        } else {
            if (drydfg == 1) slds = accsdp + slds * (1.0 - remsdp);
            drydfg = prec > 0.0 ? 0 : 1;
        }
        out->SOSLD[t] = sosld;
        out->SLDS[t] = slds;
    }
}

/* ============================================================================================== */
/* IWTGAS  (IWTGAS.py:65-183), IMPLND; metric: entry conversions in the wrapper, SOTMP stays deg C   */
/* ============================================================================================== */
void hsp2o_iwtgas(const iwtgas_ui *ui, const iwtgas_in *in, const iwtgas_out *out, int64_t nsteps) {
    const double EFACTA = 407960., MFACTA = 0.2266;
    const double slifac = ui->SLIFAC;
    double sotmp = ui->SOTMP, sodox = ui->SODOX, soco2 = ui->SOCO2;
    const double elevgc = pow((288.0 - 0.00198 * ui->ELEV) / 288.0, 5.256);
    const int metric = ui->uunits == 2.0;
    double awtf = in->AWTF[0], bwtf = in->BWTF[0]; /* raw until the first DAYFG interval (IWTGAS.py:119-120) */
    for (int64_t t = 0; t < nsteps; ++t) {
        const double airtc = (in->AIRTMP[t] - 32.0) * 0.555;
        const double suro = in->SURO[t], wyield = in->WYIELD[t], surli = in->SURLI[t];
        const double slitmp = in->SLITMP[t], slidox = in->SLIDOX[t], slico2 = in->SLICO2[t];
        if ((int)in->DAYFG[t] == 1) {
            awtf = (in->AWTF[t] - 32.0) * 0.555;
            bwtf = in->BWTF[t];
        }
        if (suro > 0.0) {
            sotmp = awtf + bwtf * airtc;
            if (sotmp < 0.5) sotmp = 0.5;
            if (wyield > 0.0) sotmp = 0.5;
            double dummy = sotmp * (0.007991 - 0.77774e-4 * sotmp);
            sodox = (14.652 + sotmp * (-0.41022 + dummy)) * elevgc;
            const double abstmp = sotmp + 273.16;
            dummy = 2385.73 / abstmp - 14.0184 + 0.0152642 * abstmp;
            soco2 = pow(10.0, dummy) * 3.16e-04 * elevgc * 12000.0;
            if (surli > 0.0 && slifac > 0.0 && slitmp >= -1.0e10) { /* IWTGAS.py:151-158 */
                sotmp = slitmp * slifac + sotmp * (1.0 - slifac);
                if (slidox >= 0.0) sodox = slidox * slifac + sodox * (1.0 - slifac);
                if (slico2 >= 0.0) soco2 = slico2 * slifac + soco2 * (1.0 - slifac);
            }
            out->SOHT[t] = sotmp * suro * EFACTA;
            out->SODOXM[t] = sodox * suro * MFACTA;
            out->SOCO2M[t] = soco2 * suro * MFACTA;
            out->SOTMP[t] = metric ? sotmp : (sotmp * 9.0 / 5.0) + 32.0; /* IWTGAS.py:160-163 */
            out->SODOX[t] = sodox;
            out->SOCO2[t] = soco2;
        } else {
            sotmp = -1.0e30;
            sodox = -1.0e30;
            soco2 = -1.0e30;
            out->SOHT[t] = 0.0;
            out->SODOXM[t] = 0.0;
            out->SOCO2M[t] = 0.0;
            out->SOTMP[t] = sotmp;
            out->SODOX[t] = sodox;
            out->SOCO2[t] = soco2;
        }
        if (metric) { /* IWTGAS.py:179-182 */
            out->SOHT[t] = out->SOHT[t] / 3.96567 * 2.471;
            out->SODOXM[t] = out->SODOXM[t] / 2.205 * 2.471;
            out->SOCO2M[t] = out->SOCO2M[t] / 2.205 * 2.471;
        }
    }
}

/* ============================================================================================== */
/* PQUAL  (PQUAL.py:147-397), one constituent, English units                                       */
/* ============================================================================================== */
void hsp2o_pqual(const pqual_ui *ui, const pqual_in *in, const pqual_out *out, int64_t nsteps) {
    const double delt60 = ui->delt60, slifac = ui->SLIFAC, ilifac = ui->ILIFAC, alifac = ui->ALIFAC;
    const double QSDFG = ui->QSDFG, QSOFG = ui->QSOFG, QIFWFG = ui->QIFWFG, QAGWFG = ui->QAGWFG;
    double sqo = QSOFG != 0.0 ? ui->SQO : 0.0; /* PQUAL.py:161-164 */
    const double wsfac = 2.30 / ui->WSQOP;
    double soqo = 0.0, soqs = 0.0;
    for (int64_t t = 0; t < nsteps; ++t) {
        const int dayfg = (int)in->DAYFG[t];
        const double suro = in->SURO[t], ifwo = in->IFWO[t], agwo = in->AGWO[t], pero = in->PERO[t];
        const double wssd = in->WSSD[t], scrsd = in->SCRSD[t], sliqsp = in->SLIQSP[t];
        const double sliqo = 0.0; /* the name is rebound to a zeros output array (PQUAL.py:205,222) */
        const double potfw = in->POTFW[t], potfs = in->POTFS[t], acqop = in->ACQOP[t], remqop = in->REMQOP[t];
        const double iliqc = in->ILIQC[t], aliqc = in->ALIQC[t];
        (void)dayfg;
        double suroqs = 0.0, soqsp = -1.0e30, scrqs = 0.0, washqs = 0.0;
        if (QSDFG != 0.0) { /* qualsd, PQUAL.py:250-276 */
            if (wssd == 0.0)
                washqs = 0.0;
            else if (sliqsp >= 0.0)
                washqs = wssd * (sliqsp * slifac + potfw * (1.0 - slifac));
            else
                washqs = wssd * potfw;
            scrqs = scrsd == 0.0 ? 0.0 : scrsd * potfs;
            soqs = washqs + scrqs;
            const double lsosed = wssd + scrsd;
            soqsp = lsosed > 0.0 ? soqs / lsosed : -1.0e30;
            suroqs = soqs;
        }
        double suroqo = 0.0, adtot = 0.0, adfxfx = 0.0, adcnfx = 0.0, soqoc = -1.0e30;
        if (QSOFG != 0.0) { /* qualof, PQUAL.py:284-325 */
            if (dayfg && QSOFG == 1) sqo = acqop + sqo * (1.0 - remqop);
            adfxfx = in->PQADFX[t];
            adcnfx = in->PQADCN[t] * in->PREC[t] * 3630.0;
            adtot = adfxfx + adcnfx;
            const double intot = adtot + sliqo;
            if (QSOFG == 2) {
                double dummy = remqop + intot / (acqop / remqop);
                if (dummy > 1.0) dummy = 1.0;
                sqo = acqop * (delt60 / 24.0) + sqo * pow(1.0 - dummy, delt60 / 24.0);
            }
            sqo = sqo + intot;
            soqo = 0.0;
            if (suro > 0.0 && sqo > 0.0) {
                double dummy = suro * wsfac;
                if (dummy < 1.0e-5)
                    soqo = 0.0;
                else {
                    dummy = 1.0 - exp(-dummy);
                    soqo = sqo * dummy;
                    sqo = sqo - soqo;
                }
            }
            soqoc = suro > 0.0 ? soqo / suro : -1.0e30;
            suroqo = soqo;
        }
        const double soqual = suroqs + suroqo;
        double poqual = soqual, ioqual = 0.0, aoqual = 0.0;
        const double soqc = suro > 0.0 ? soqual / suro : -1.0e30;
        if (QIFWFG != 0.0) { /* qualif, PQUAL.py:339-353 */
            double ioqc = in->IOQCP[t] * 3630.0;
            if (ui->VIQCFG == 3 || ui->VIQCFG == 4) ioqc = ioqc * 6.238e-5;
            if (ifwo > 0.0) {
                const double ioqce = iliqc >= 0.0 ? iliqc * ilifac + ioqc * (1.0 - ilifac) : ioqc;
                ioqual = ioqce * ifwo;
            } else
                ioqual = 0.0;
            poqual = poqual + ioqual;
        }
        if (QAGWFG != 0.0) { /* qualgw, PQUAL.py:356-371 */
            double aoqc = in->AOQCP[t] * 3630.0;
            if (ui->VAQCFG == 3 || ui->VAQCFG == 4) aoqc = aoqc * 6.238e-5;
            if (agwo > 0.0) {
                const double aoqce = aliqc >= 0.0 ? aliqc * alifac + aoqc * (1.0 - alifac) : aoqc;
                aoqual = aoqce * agwo;
            } else
                aoqual = 0.0;
            poqual = poqual + aoqual;
        }
        const double poqc = pero > 0.0 ? poqual / pero : -1.0e30;
        out->SOQUAL[t] = soqual;
        out->IOQUAL[t] = ioqual;
        out->AOQUAL[t] = aoqual;
        out->POQUAL[t] = poqual;
        out->SQO[t] = sqo;
        out->SOQSP[t] = soqsp;
        out->SOQOC[t] = soqoc > -1 ? soqoc / 3630.0 : soqoc;
        out->SOQC[t] = soqc / 3630.0;
        out->IOQC[t] = ifwo > 0.0 ? ioqual / ifwo / 3630.0 : -1.0e30;
        out->AOQC[t] = agwo > 0.0 ? aoqual / agwo / 3630.0 : -1.0e30;
        out->POQC[t] = pero > 0.0 ? poqc / 3630.0 : -1.0e30;
        out->WASHQS[t] = washqs;
        out->SCRQS[t] = scrqs;
        out->SOQS[t] = soqs;
        out->SOQO[t] = soqo;
        out->PQADWT[t] = adcnfx;
        out->PQADDR[t] = adfxfx;
        out->PQADEP[t] = adtot;
        out->SLIQO[t] = 0.0;  /* allocated, never written (PQUAL.py:205-206) */
        out->INFLOW[t] = 0.0;
    }
}

/* ============================================================================================== */
/* IQUAL  (IQUAL.py:131-284), one constituent, English units                                       */
/* ============================================================================================== */
void hsp2o_iqual(const iqual_ui *ui, const iqual_in *in, const iqual_out *out, int64_t nsteps) {
    const double delt60 = ui->delt60, slifac = ui->SLIFAC, QSDFG = ui->QSDFG, QSOFG = ui->QSOFG;
    double sqo = ui->SQO;
    const double wsfac = 2.30 / ui->WSQOP;
    double soqo = 0.0, remqop = 0.0, soqs = 0.0, soqoc = 0.0, soqsp = 0.0; /* carried (IQUAL.py:184-188) */
    for (int64_t t = 0; t < nsteps; ++t) {
        const double suro = in->SURO[t], sosld = in->SOSLD[t], sliqsp = in->SLIQSP[t];
        const int dayfg = (int)in->DAYFG[t];
        const double sliqo = 0.0; /* rebound to a zeros output array (IQUAL.py:174) */
        const double potfw = in->POTFW[t];
        double acqop = in->ACQOP[t];
        double suroqs = 0.0;
        if (QSDFG != 0.0) { /* washsd, IQUAL.py:201-217 */
            if (sosld == 0.0)
                soqs = 0.0;
            else if (sliqsp >= 0.0) {
                soqsp = sliqsp * slifac + potfw * (1.0 - slifac);
                soqs = sosld * soqsp;
            } else {
                soqsp = potfw;
                soqs = sosld * potfw;
            }
            suroqs = soqs;
        }
        double suroqo = 0.0, adtot = 0.0, adfxfx = 0.0, adcnfx = 0.0;
        if (QSOFG != 0.0) {
            if (QSOFG >= 1) { /* washof, IQUAL.py:226-257 */
                if (dayfg == 1) {
                    remqop = in->REMQOP[t];
                    if (QSOFG == 1) sqo = acqop + sqo * (1.0 - remqop);
                }
                adfxfx = in->IQADFX[t];
                adcnfx = in->IQADCN[t] * in->PREC[t] * 3630.0;
                adtot = adfxfx + adcnfx;
                if (QSOFG == 2) {
                    double dummy = remqop + (adtot + sliqo) / (acqop / remqop);
                    if (dummy > 1.0) dummy = 1.0;
                    sqo = acqop * (delt60 / 24.0) + sqo * pow(1.0 - dummy, delt60 / 24.0);
                }
                sqo = sqo + sliqo + adtot;
                soqo = 0.0;
                if (suro > 0.0 && sqo > 0.0) {
                    soqo = sqo * (1.0 - exp(-suro * wsfac));
                    sqo = sqo - soqo;
                }
                soqoc = suro > 0.0 ? soqo / suro / 3630.0 : -1.0e30;
            } else if (QSOFG == -1) { /* IQUAL.py:259-269 */
                acqop = acqop * 0.2266;
                soqo = suro * acqop;
                soqoc = acqop / 3630.0;
                sqo = 0.0;
            }
            suroqo = soqo;
        }
        const double soqual = suroqs + suroqo;
        out->SOQUAL[t] = soqual;
        out->SOQC[t] = suro > 0.0 ? soqual / suro / 3630.0 : -1.0e30;
        out->SQO[t] = sqo;
        out->SOQS[t] = soqs;
        out->SOQOC[t] = soqoc;
        out->SOQO[t] = soqo;
        out->SOQSP[t] = soqsp;
        out->IQADWT[t] = adcnfx;
        out->IQADDR[t] = adfxfx;
        out->IQADEP[t] = adtot;
        out->SLIQO[t] = 0.0;
        out->INFLOW[t] = 0.0;
    }
}

/* ============================================================================================== */
/* Multi-threaded CPU baseline: SNOW -> PWATER per segment, segments in parallel (OpenMP)          */
/* ============================================================================================== */
int hsp2o_perlnd_batch(int64_t nseg, int64_t nsteps, double delt, const double *forc, const double *par,
                       const double *mon, const double *cal, double *sums, int64_t *errors, int nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#else
    (void)nthreads;
#endif
    const int64_t n1 = nsteps + 1;
#pragma omp parallel num_threads(used)
    {
        /* per-thread working set: constant-parameter series + the 22 + 31 output series */
        enum { NCONST = 18, NOUT = 22 + 31 };
        double *buf = (double *)malloc(sizeof(double) * (size_t)nsteps * (NCONST + NOUT));
        double *cst[NCONST];
        double *o[NOUT];
        for (int i = 0; i < NCONST; ++i) cst[i] = buf + (size_t)i * nsteps;
        for (int i = 0; i < NOUT; ++i) o[i] = buf + (size_t)(NCONST + i) * nsteps;
        double *zeros = cst[16];
        for (int64_t t = 0; t < nsteps; ++t) zeros[t] = 0.0;

#pragma omp for schedule(dynamic, 4)
        for (int64_t s = 0; s < nseg; ++s) {
            const double *p = par + s * HSP2O_BATCH_NPAR;
            const double *f = forc + s * 6 * nsteps;
            const double *m = mon + s * 4 * n1;
            /* the reference wrappers broadcast constant parameters to per-step arrays (SNOW.py:49-51,
               PWATER.py:51-53); do the same so the work is the same */
            static const int bc[16] = {BP_CCFACT, BP_COVIND, BP_MGMELT, BP_MWATER, BP_SHADE, BP_SNOEVP, BP_SNOWCF,
                                       BP_KMELT, BP_AGWRC, BP_DEEPFR, BP_INFILT, BP_KVARY, BP_LZSN, BP_PETMIN,
                                       BP_PETMAX, BP_INTFW};
            for (int i = 0; i < 16; ++i) {
                double v = p[bc[i]];
                double *d = cst[i];
                for (int64_t t = 0; t < nsteps; ++t) d[t] = v;
            }
            double *irc = cst[17];
            for (int64_t t = 0; t < nsteps; ++t) irc[t] = p[BP_IRC];

            snow_ui su = {delt, 0.0, p[BP_ICEFG], p[BP_SNOPFG], p[BP_MELEV], p[BP_PACKF], p[BP_PACKI], p[BP_PACKW],
                          p[BP_PAKTMP], p[BP_RDCSN], p[BP_RDENPF], p[BP_SKYCLR], p[BP_TBASE], p[BP_TSNOW],
                          p[BP_COVINX], p[BP_DULL], p[BP_XLNMLT], 1.0};
            snow_in si;
            si.AIRTMP = f + 2 * nsteps; si.CCFACT = cst[0]; si.CLOUD = zeros; si.COVIND = cst[1];
            si.DTMPG = f + 5 * nsteps; si.HR6IND = cal + 2 * n1; si.HRFG = cal; si.KMELT = cst[7];
            si.MGMELT = cst[2]; si.MWATER = cst[3]; si.PREC = f; si.SEASONS = cal + 3 * n1; si.SHADE = cst[4];
            si.SNOEVP = cst[5]; si.SNOWCF = cst[6]; si.SOLRAD = f + 4 * nsteps; si.WINMOV = f + 3 * nsteps;
            snow_out so;
            {
                double **q = (double **)&so;
                for (int i = 0; i < 22; ++i) q[i] = o[i];
            }
            int64_t e1[1], e2[10];
            hsp2o_snow(&su, &si, &so, nsteps, e1);

            pwater_ui pu = {delt, p[BP_ICEFG], 1.0, p[BP_IFFCFG], p[BP_IFRDFG], p[BP_RTOPFG], p[BP_UZFG], p[BP_VLEFG],
                            p[BP_AGWETP], p[BP_AGWS], p[BP_BASETP], p[BP_CEPS], p[BP_GWVS], p[BP_IFWS], p[BP_INFEXP],
                            p[BP_INFILD], p[BP_LSUR], p[BP_LZS], p[BP_SLSUR], p[BP_SURS], p[BP_UZS], p[BP_FOREST],
                            p[BP_FZG], p[BP_FZGL], 1.0};
            pwater_in pi;
            pi.AGWLI = zeros; pi.AGWRC = cst[8]; pi.AIRTMP = f + 2 * nsteps; pi.CEPSC = m; pi.DEEPFR = cst[9];
            pi.IFWLI = zeros; pi.INFILT = cst[10]; pi.INTFW = cst[15]; pi.IRC = irc; pi.KVARY = cst[11];
            pi.LGTMP = zeros; pi.LZETP = m + 3 * n1; pi.LZLI = zeros; pi.LZSN = cst[12]; pi.NSUR = m + 2 * n1;
            pi.PACKI = so.PACKI; pi.PETINP = f + nsteps; pi.PETMAX = cst[14]; pi.PETMIN = cst[13]; pi.PREC = f;
            pi.RAINF = so.RAINF; pi.SNOCOV = so.SNOCOV; pi.SURLI = zeros; pi.UZLI = zeros; pi.UZSN = m + n1;
            pi.WYIELD = so.WYIELD; pi.DAYFG = cal + n1; pi.HRFG = cal;
            pwater_out po;
            {
                double **q = (double **)&po;
                for (int i = 0; i < 31; ++i) q[i] = o[22 + i];
            }
            hsp2o_pwater(&pu, &pi, &po, nsteps, e2);

            /* checksum of the default SAVE set (hsp2tools/data/SaveTable.csv): 13 SNOW + 11 PWATER series */
            const double *sv[24] = {so.COVINX, so.DULL, so.PACKF, so.PACKI, so.PACKW, so.PACK, so.PAKTMP, so.RAINF,
                                    so.RDENPF, so.SKYCLR, so.SNOCOV, so.WYIELD, so.XLNMLT, po.AGWO, po.AGWS,
                                    po.CEPS, po.GWVS, po.IFWO, po.IFWS, po.LZS, po.PERO, po.SURO, po.SURS, po.UZS};
            for (int v = 0; v < 24; ++v) {
                double acc = 0.0;
                const double *a = sv[v];
                for (int64_t t = 0; t < nsteps; ++t) acc += a[t];
                sums[s * 24 + v] = acc;
            }
            if (errors) {
                errors[s * 11] = e1[0];
                for (int i = 0; i < 10; ++i) errors[s * 11 + 1 + i] = e2[i];
            }
        }
        free(buf);
    }
    return used;
}
