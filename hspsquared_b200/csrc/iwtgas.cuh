// iwtgas.cuh -- one IWTGAS interval (temperature, dissolved oxygen and CO2 of impervious-land runoff) for one segment.
//
// Behaviour: HSPF HIMPGAS as HSP2 runs it, src/hsp2/hsp2/IWTGAS.py:122-183 (_iwtgas_ loop body).  Metric runs: AWTF goes to deg F
// first (:112), wyield / suro arrive as `x * 0.0394` of the stored series, SOTMP is stored in deg C (:160-163), fluxes by save().
// Differences from PWTGAS' surface branch that matter for parity: the runoff temperature is a regression on air
// temperature (AWTF + BWTF * airtc, coefficients refreshed on DAYFG intervals), the literal 0.555, and the lateral
// inflow block: DO and CO2 are only mixed when the lateral TEMPERATURE is present (they nest under it, :151-158).
#pragma once
#include "land_common.cuh"

struct IwtgasState {
    double awtf, bwtf;  // the day's regression coefficients (deg C, -)
    template <class SEG>
    __device__ __forceinline__ void load(const SEG &s) {
        awtf = s.st(S_IWG_AWTF);
        bwtf = s.st(S_IWG_BWTF);
    }
    template <class SEG>
    __device__ __forceinline__ void store(const SEG &s) const {
        s.st(S_IWG_AWTF, awtf);
        s.st(S_IWG_BWTF, bwtf);
    }
};

template <class SEG>
__device__ __forceinline__ void iwtgas_step(const SEG &s, IwtgasState &w, double airtmp, double wyield, double suro) {
    const double EFACTA = 407960., MFACTA = 0.2266;
    const double airtc = (airtmp - 32.0) * 0.555;
    if (s.calfl & HSP2G_CAL_DAYFG) {  // IWTGAS.py:132-134
        double awtf = s.daily(FB_M_AWTF, M_AWTF, P_AWTF);
        if (s.metric()) awtf = (awtf * 9. / 5.) + 32.;  // IWTGAS.py:112
        w.awtf = (awtf - 32.0) * 0.555;
        w.bwtf = s.daily(FB_M_BWTF, M_BWTF, P_BWTF);
    }
    double sotmp_out = -1.0e30, sodox = -1.0e30, soco2 = -1.0e30, soht = 0.0, sodoxm = 0.0, soco2m = 0.0;
    if (suro > 0.0) {
        const double elevgc = s.par(P_ELEVGC);
        double sotmp = w.awtf + w.bwtf * airtc;
        if (sotmp < 0.5) sotmp = 0.5;
        if (wyield > 0.0) sotmp = 0.5;
        double dummy = sotmp * (0.007991 - 0.77774e-4 * sotmp);
        sodox = (14.652 + sotmp * (-0.41022 + dummy)) * elevgc;
        const double abstmp = sotmp + 273.16;
        dummy = HDIV(2385.73, abstmp) - 14.0184 + 0.0152642 * abstmp;
        soco2 = hsp_pow(10.0, dummy) * 3.16e-04 * elevgc * 12000.0;
        const double slifac = s.par(P_SLIFAC);
        const double slitmp = s.forcing(F_SLITMP);
        if (s.forcing(F_SURLI) > 0.0 && slifac > 0.0 && slitmp >= -1.0e10) {
            sotmp = slitmp * slifac + sotmp * (1.0 - slifac);
            const double slidox = s.forcing(F_SLIDOX), slico2 = s.forcing(F_SLICO2);
            if (slidox >= 0.0) sodox = slidox * slifac + sodox * (1.0 - slifac);
            if (slico2 >= 0.0) soco2 = slico2 * slifac + soco2 * (1.0 - slifac);
        }
        soht = sotmp * suro * EFACTA;
        sodoxm = sodox * suro * MFACTA;
        soco2m = soco2 * suro * MFACTA;
        sotmp_out = s.metric() ? sotmp : (sotmp * 9.0 / 5.0) + 32.0;
    }
    s.save(O_IWTGAS_SOTMP, sotmp_out);
    s.save(O_IWTGAS_SODOX, sodox);
    s.save(O_IWTGAS_SOCO2, soco2);
    s.save(O_IWTGAS_SOHT, soht);
    s.save(O_IWTGAS_SODOXM, sodoxm);
    s.save(O_IWTGAS_SOCO2M, soco2m);
}
