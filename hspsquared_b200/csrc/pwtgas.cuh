// pwtgas.cuh -- one PWTGAS interval (outflow water temperature, dissolved oxygen, CO2) for one segment.
//
// Behaviour: HSPF HPERGAS as HSP2 runs it, src/hsp2/hsp2/PWTGAS.py:172-334 (_pwtgas_ loop body).  Metric runs: the caller hands
// the soil temperatures in deg F and the flows in inches as the reference's entry conversions make them (:167-174,
// land_kernels.cuh); the three outflow temperatures are stored in deg C (:310-320), heat and mass fluxes converted by save().
// Stateless: every quantity is rebuilt each interval; -1e30 marks "no flow" and is multiplied by the zero flow in
// the flux terms exactly as the reference does.
#pragma once
#include "land_common.cuh"
#include "pstemp.cuh"

__device__ __forceinline__ void mix_lateral(double &t, double &ox, double &co2, double lit, double lio, double lic,
                                            double fac) {
    if (lit >= -1.0e10) t = lit * fac + t * (1.0 - fac);
    if (lio >= 0.0) ox = lio * fac + ox * (1.0 - fac);
    if (lic >= 0.0) co2 = lic * fac + co2 * (1.0 - fac);
}

// interflow / groundwater DO and CO2 concentrations (optionally monthly), interpolated once per simulated day
struct PwtgasDaily {
    ColdD idoxp, ico2p, adoxp, aco2p;
    __device__ __forceinline__ void bind(double *c) {
        ColdD *m[NCOLD_PWTGAS] = {&idoxp, &ico2p, &adoxp, &aco2p};
        for (int k = 0; k < NCOLD_PWTGAS; ++k) m[k]->p = c + k * LAND_TILE;
    }
    template <class SEG>
    __device__ __forceinline__ void refresh_daily(const SEG &s) {
        const int fb[4] = {FB_M_IDOXP, FB_M_ICO2P, FB_M_ADOXP, FB_M_ACO2P};
        const int mid[4] = {M_IDOXP, M_ICO2P, M_ADOXP, M_ACO2P};
        const int pid[4] = {P_IDOXP, P_ICO2P, P_ADOXP, P_ACO2P};
        double v[4];
        s.daily_batch(fb, mid, pid, v);
        idoxp = v[0]; ico2p = v[1]; adoxp = v[2]; aco2p = v[3];
    }
};

template <class SEG>
__device__ __forceinline__ void pwtgas_step(const SEG &s, const PwtgasDaily &g, const SoilTemps &soil, double wyield, double suro,
                                            double ifwo, double agwo) {
    const double EFACTA = 407960., MFACTA = 0.2266;
    const double elevgc = s.par(P_ELEVGC);
    const double surli = s.forcing(F_SURLI), ifwli = s.forcing(F_IFWLI), agwli = s.forcing(F_AGWLI);

    double sotmp = -1.0e30, sodox = -1.0e30, soco2 = -1.0e30;
    if (suro > 0.0) {
        sotmp = (soil.sltmp - 32.0) * 5. / 9.;
        if (sotmp < 0.5) sotmp = 0.5;
        if (s.flag(FB_PWG_CSNO)) {
            if (wyield > 0.0) sotmp = 0.5;
        }
        double dummy = sotmp * (0.007991 - 0.77774E-4 * sotmp);
        sodox = (14.652 + sotmp * (-0.41022 + dummy)) * elevgc;
        const double abstmp = sotmp + 273.16;
        dummy = 2385.73 / abstmp - 14.0184 + 0.0152642 * abstmp;
        soco2 = hsp_pow(10.0, dummy) * 3.16e-04 * elevgc * 12000.0;
        const double slifac = s.par(P_SLIFAC);
        if (surli > 0.0 && slifac > 0.0)
            mix_lateral(sotmp, sodox, soco2, s.forcing(F_SLITMP), s.forcing(F_SLIDOX), s.forcing(F_SLICO2), slifac);
    }
    double iotmp = -1.0e30, iodox = -1.0e30, ioco2 = -1.0e30;
    if (ifwo > 0.0) {
        iotmp = (soil.ultmp - 32.0) * 5. / 9.;
        if (iotmp < 0.5) iotmp = 0.5;
        iodox = g.idoxp;
        ioco2 = g.ico2p;
        const double ilifac = s.par(P_ILIFAC);
        if (ifwli > 0.0 && ilifac > 0.0)
            mix_lateral(iotmp, iodox, ioco2, s.forcing(F_ILITMP), s.forcing(F_ILIDOX), s.forcing(F_ILICO2), ilifac);
    }
    double aotmp = -1.0e30, aodox = -1.0e30, aoco2 = -1.0e30;
    if (agwo > 0.0) {
        aotmp = (soil.lgtmp - 32.0) * 5. / 9.;
        if (aotmp < 0.5) aotmp = 0.5;
        aodox = g.adoxp;
        aoco2 = g.aco2p;
        const double alifac = s.par(P_ALIFAC);
        if (agwli > 0.0 && alifac > 0.0)
            mix_lateral(aotmp, aodox, aoco2, s.forcing(F_ALITMP), s.forcing(F_ALIDOX), s.forcing(F_ALICO2), alifac);
    }
    const double soht = sotmp * suro * EFACTA, ioht = iotmp * ifwo * EFACTA, aoht = aotmp * agwo * EFACTA;
    const double sodoxm = sodox * suro * MFACTA, iodoxm = iodox * ifwo * MFACTA, aodoxm = aodox * agwo * MFACTA;
    const double soco2m = soco2 * suro * MFACTA, ioco2m = ioco2 * ifwo * MFACTA, aoco2m = aoco2 * agwo * MFACTA;

    s.save(O_PWTGAS_POHT, soht + ioht + aoht);
    s.save(O_PWTGAS_PODOXM, sodoxm + iodoxm + aodoxm);
    s.save(O_PWTGAS_POCO2M, soco2m + ioco2m + aoco2m);
    const bool degf = !s.metric();
    s.save(O_PWTGAS_SOTMP, (sotmp > -1e28 && degf) ? (sotmp * 9.0 / 5.0) + 32.0 : sotmp);
    s.save(O_PWTGAS_IOTMP, (iotmp > -1e28 && degf) ? (iotmp * 9.0 / 5.0) + 32.0 : iotmp);
    s.save(O_PWTGAS_AOTMP, (aotmp > -1e28 && degf) ? (aotmp * 9.0 / 5.0) + 32.0 : aotmp);
    s.save(O_PWTGAS_SODOX, sodox);
    s.save(O_PWTGAS_SOCO2, soco2);
    s.save(O_PWTGAS_IODOX, iodox);
    s.save(O_PWTGAS_IOCO2, ioco2);
    s.save(O_PWTGAS_AODOX, aodox);
    s.save(O_PWTGAS_AOCO2, aoco2);
    s.save(O_PWTGAS_SOHT, soht);
    s.save(O_PWTGAS_IOHT, ioht);
    s.save(O_PWTGAS_AOHT, aoht);
    s.save(O_PWTGAS_SODOXM, sodoxm);
    s.save(O_PWTGAS_SOCO2M, soco2m);
    s.save(O_PWTGAS_IODOXM, iodoxm);
    s.save(O_PWTGAS_IOCO2M, ioco2m);
    s.save(O_PWTGAS_AODOXM, aodoxm);
    s.save(O_PWTGAS_AOCO2M, aoco2m);
}
