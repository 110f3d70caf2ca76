"""CPU restatement of the colormath 3.0.0 pipeline the reference calls.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  PARITY UNPINNED: colormath is a
third-party dependency (requirements.txt:6, pinned 3.0.0) that is neither
installed nor vendored, so this file restates its *published* algorithm and is
anchored on
  * the reference's call sites: make_data_tables.py:55-70 (convert_color to
    LabColor, delta_e_cie2000, ``int()`` truncation) and palette.py:6-15
    (``sRGBColor(r, g, b, is_upscaled=True)``);
  * the CIEDE2000 test data published by Sharma, Wu & Dalal (2005), Table 1
    (tests/golden/ciede2000_sharma.json) -- every pair the colormath variant
    agrees with the paper on;
  * the known answers recorded in SURVEY.md Appendix C.

colormath semantics restated (numpy float64 throughout, like colormath):
  sRGB -> linear:   v/255; v <= 0.04045 ? v/12.92 : ((v+0.055)/1.055)**2.4
  linear -> XYZ:    row vector times the sRGB matrix below, clamped at 0;
                    native illuminant D65, no chromatic adaptation
  XYZ -> Lab:       D65 / 2-degree white (0.95047, 1.0, 1.08883),
                    eps = 216/24389, linear branch 7.787*t + 16/116
  delta_e_cie2000:  Kl = Kc = Kh = 1; colormath's vectorised variant:
                    avg_Hp = ((|h1-h2| > 180)*360 + h1 + h2)/2  (no -360 branch,
                    no zero-chroma special case);
                    dhp = (h2-h1) + (|h2-h1| > 180)*360 - (h2 > h1)*720
"""

import numpy as np

_RGB_TO_XYZ = np.array([
    [0.412424, 0.212656, 0.0193324],
    [0.357579, 0.715158, 0.119193],
    [0.180464, 0.0721856, 0.950444],
], dtype=np.float64)
_WHITE_D65_2 = np.array([0.95047, 1.00000, 1.08883], dtype=np.float64)
_CIE_E = 216.0 / 24389.0


def srgb255_to_lab(rgb) -> np.ndarray:
    v = np.asarray(rgb, dtype=np.float64) / 255.0
    lin = np.where(v <= 0.04045, v / 12.92, np.power((v + 0.055) / 1.055, 2.4))
    xyz = np.maximum(np.dot(lin, _RGB_TO_XYZ), 0.0)
    t = xyz / _WHITE_D65_2
    f = np.where(t > _CIE_E, np.power(t, 1.0 / 3.0), 7.787 * t + 16.0 / 116.0)
    return np.array([
        116.0 * f[1] - 16.0, 500.0 * (f[0] - f[1]), 200.0 * (f[1] - f[2])],
        dtype=np.float64)


def delta_e_cie2000(lab1, lab2) -> float:
    L1, a1, b1 = (np.float64(x) for x in lab1)
    L2, a2, b2 = (np.float64(x) for x in lab2)
    avg_Lp = (L1 + L2) / 2.0
    C1 = np.sqrt(a1 ** 2 + b1 ** 2)
    C2 = np.sqrt(a2 ** 2 + b2 ** 2)
    avg_C = (C1 + C2) / 2.0
    G = 0.5 * (1 - np.sqrt(
        np.power(avg_C, 7.0) / (np.power(avg_C, 7.0) + np.power(25.0, 7.0))))
    a1p = (1.0 + G) * a1
    a2p = (1.0 + G) * a2
    C1p = np.sqrt(a1p ** 2 + b1 ** 2)
    C2p = np.sqrt(a2p ** 2 + b2 ** 2)
    avg_Cp = (C1p + C2p) / 2.0
    h1p = np.degrees(np.arctan2(b1, a1p))
    h1p += (h1p < 0) * 360
    h2p = np.degrees(np.arctan2(b2, a2p))
    h2p += (h2p < 0) * 360
    avg_Hp = (((np.fabs(h1p - h2p) > 180) * 360) + h1p + h2p) / 2.0
    T = (1 - 0.17 * np.cos(np.radians(avg_Hp - 30))
         + 0.24 * np.cos(np.radians(2 * avg_Hp))
         + 0.32 * np.cos(np.radians(3 * avg_Hp + 6))
         - 0.2 * np.cos(np.radians(4 * avg_Hp - 63)))
    dh = h2p - h1p
    delta_hp = dh + (np.fabs(dh) > 180) * 360
    delta_hp -= (h2p > h1p) * 720
    delta_Lp = L2 - L1
    delta_Cp = C2p - C1p
    delta_Hp = 2 * np.sqrt(C2p * C1p) * np.sin(np.radians(delta_hp) / 2.0)
    S_L = 1 + ((0.015 * np.power(avg_Lp - 50, 2))
               / np.sqrt(20 + np.power(avg_Lp - 50, 2.0)))
    S_C = 1 + 0.045 * avg_Cp
    S_H = 1 + 0.015 * avg_Cp * T
    delta_ro = 30 * np.exp(-(np.power(((avg_Hp - 275) / 25), 2.0)))
    R_C = np.sqrt(
        np.power(avg_Cp, 7.0) / (np.power(avg_Cp, 7.0) + np.power(25.0, 7.0)))
    R_T = -2 * R_C * np.sin(2 * np.radians(delta_ro))
    return float(np.sqrt(
        np.power(delta_Lp / S_L, 2) + np.power(delta_Cp / S_C, 2)
        + np.power(delta_Hp / S_H, 2)
        + R_T * (delta_Cp / S_C) * (delta_Hp / S_H)))


def diff_matrix(rgb16) -> np.ndarray:
    """make_data_tables.py:55-70: int()-truncated 16x16 dE2000 matrix.

    rgb16[c] is the sRGB triple of the colour whose HGRColours value is c.
    """
    labs = [srgb255_to_lab(rgb16[c]) for c in range(16)]
    dm = np.zeros((16, 16), dtype=np.int32)
    for i in range(16):
        for j in range(16):
            dm[i, j] = int(delta_e_cie2000(labs[i], labs[j]))
    return dm


def diff_matrix_float(rgb16) -> np.ndarray:
    labs = [srgb255_to_lab(rgb16[c]) for c in range(16)]
    return np.array([[delta_e_cie2000(labs[i], labs[j]) for j in range(16)]
                     for i in range(16)], dtype=np.float64)
