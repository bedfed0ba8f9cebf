"""Chesapeake-Bay-style full-chain PERLND segment (ATEMP, SNOW, PWATER RTOPFG=1, SEDMNT, PSTEMP, PWTGAS) as UCI tables:
the numbers of the reference's tests/land_spec/hwmA51800.uci:73-262 on top of the test10 PERLND defaults.  Base segment
of BASELINE.json configs[3] (250k-segment full chain) and of the `cbp_*` parity cases (oracle/cases.py)."""
from . import test10
from .test10 import monthly


def perlnd():
    """Full chain, Chesapeake-Bay style land segment (tests/land_spec/hwmA51800.uci:73-262)."""
    t = test10.perlnd()
    t["ACTIVITY"] = {"ATEMP": 1, "SNOW": 1, "PWATER": 1, "SEDMNT": 1, "PSTEMP": 1, "PWTGAS": 1}
    t["ATEMP"] = {"PARAMETERS": {"ELDAT": 150.0, "AIRTMP": 23.1476}}
    t["SNOW"]["PARAMETERS"].update({"LAT": 36.69108, "MELEV": 55.78, "SHADE": 0.05, "SNOWCF": 1.3, "COVIND": 1.08,
                                    "SNOEVP": 0.13, "MWATER": 0.03, "MGMELT": 0.03})
    t["SNOW"]["STATES"].update({"PACKF": 0.0, "PACKI": 0.0, "PACKW": 0.0, "RDENPF": 0.15, "DULL": 100.0,
                                "PAKTMP": 30.0, "COVINX": 1.08, "SKYCLR": 0.9})
    p = t["PWATER"]["PARAMETERS"]
    p.update({"RTOPFG": 1, "FOREST": 0.0, "INFILT": 0.017857, "LSUR": 461.0, "SLSUR": 0.014, "KVARY": 0.10E-05,
              "AGWRC": 0.92, "DEEPFR": 0.0, "AGWETP": 0.05, "UZSN": 0.7, "INTFW": 2.0, "IRC": 0.435477})
    t["PWATER"]["STATES"].update({"CEPS": 0.0, "UZS": 0.7, "LZS": 5.0, "AGWS": 1.0, "GWVS": 0.0})
    t["PWATER"]["MONTHLY_CEPSC"] = monthly([.067, .065, .062, .06, .064, .085, .138, .15, .146, .109, .087, .077])
    t["PWATER"]["MONTHLY_UZSN"] = monthly([.672, .672, .672, .672, .672, .784, 1.064, 1.12, 1.12, .896, .784, .728])
    t["PWATER"]["MONTHLY_NSUR"] = monthly([.2028, .2028, .1971, .2103, .2804, .3505, .3692, .3594, .3262, .2283, .2085, .2028])
    t["PWATER"]["MONTHLY_LZETP"] = monthly([.232, .221, .208, .2, .218, .312, .547, .6, .582, .416, .322, .276])
    t["SEDMNT"] = {
        "PARAMETERS": {"CRVFG": 1, "VSIVFG": 0, "SDOPFG": 1, "SMPF": 1.0, "KRER": 1.394872, "JRER": 2.0,
                       "AFFIX": 0.07675, "COVER": 0.0, "NVSI": 30.57534, "KSER": 6.974359, "JSER": 2.0,
                       "KGER": 0.0, "JGER": 4.0, "DETS": 0.0},
        "MONTHLY_COVER": monthly([.0618, .0618, .0557, .1392, .2285, .3906, .6657, .7618, .8416, .7424, .6945, .0618]),
        "MONTHLY_NVSI": monthly([.1] * 12),
    }
    t["PSTEMP"] = {
        "PARAMETERS": {"SLTVFG": 1, "ULTVFG": 1, "LGTVFG": 1, "TSOPFG": 1, "ASLT": 0.0, "BSLT": 1.0, "ULTP1": 0.0,
                       "ULTP2": 0.0, "LGTP1": 0.0, "LGTP2": 0.0, "AIRTC": 32.0, "SLTMP": 32.0, "ULTMP": 32.0,
                       "LGTMP": 50.0},
        "MONTHLY_ASLT": monthly([35.58, 37.14, 40.49, 45.09, 49.46, 53.62, 55.76, 54.6, 51.54, 46.08, 41.55, 36.97]),
        "MONTHLY_BSLT": monthly([.5] * 12),
        "MONTHLY_ULTP1": monthly([38.45, 41.25, 47.27, 55.57, 63.43, 70.92, 74.77, 72.67, 67.17, 57.34, 49.19, 40.94]),
        "MONTHLY_ULTP2": monthly([.1] * 12),
        "MONTHLY_LGTP1": monthly([46.51, 45.41, 46.57, 49.09, 52.79, 56.03, 59.01, 60.41, 59.74, 57.2, 53.27, 49.9]),
        "MONTHLY_LGTP2": monthly([0.0] * 12),
    }
    t["PWTGAS"] = {
        "PARAMETERS": {"IDVFG": 1, "ICVFG": 0, "GDVFG": 1, "GCVFG": 0, "ELEV": 55.78, "IDOXP": 0.0, "ICO2P": 0.0,
                       "ADOXP": 0.0, "ACO2P": 0.0, "SDLFAC": 0.0, "SLIFAC": 0.0, "ILIFAC": 0.0, "ALIFAC": 0.0,
                       "SOTMP": 33.0, "IOTMP": 40.0, "AOTMP": 55.0, "SODOX": 0.0, "SOCO2": 0.0, "IODOX": 0.0,
                       "IOCO2": 0.0, "AODOX": 0.0, "AOCO2": 0.0},
        "MONTHLY_IDOXP": monthly([12.7, 12.7, 11.2, 9.7, 7.4, 6.6, 6.2, 6.2, 7.5, 8.4, 9.4, 11.6]),
        "MONTHLY_ADOXP": monthly([10., 10., 10., 10., 7.5, 5.5, 5., 5., 6.5, 7.5, 9., 10.]),
    }
    return t
