"""Primer melting temperature for `swga filter` (Primers.filter_tm_range, swga/primers.py:151-165;
Primer.update_tm -> `melting.temp(seq)`, swga/workspace.py:23,147-148).

The reference takes `melting.temp` from the third-party PyPI package `melt` (requirements.txt:6, setup.py:129;
version not pinned, source NOT part of /root/reference).  **Parity unpinned**: no reference test asserts a Tm and
the package cannot be installed here, so this module restates the PUBLISHED method that package documents --
nearest-neighbour thermodynamics with the unified parameters and terminal corrections of Allawi & SantaLucia
(1997) / SantaLucia (1998), and the mono-/divalent cation correction of Owczarzy et al. (2008) -- with the
package's documented defaults (DNA 5000 nM, Na+ 10 mM, Mg2+ 20 mM, dNTPs 10 mM: the phi29 MDA reaction).
Host code, off the GPU path (at most a few thousand primers reach it).  `commands.filter(tm_fn=...)` takes any
other implementation.
"""
from math import log, sqrt

# nearest-neighbour dH (kcal/mol) and dS (cal/(K mol)), Allawi & SantaLucia (1997), Table 1
_DH = {"AA": -7.9, "TT": -7.9, "AT": -7.2, "TA": -7.2, "CA": -8.5, "TG": -8.5, "GT": -8.4, "AC": -8.4,
       "CT": -7.8, "AG": -7.8, "GA": -8.2, "TC": -8.2, "CG": -10.6, "GC": -9.8, "GG": -8.0, "CC": -8.0}
_DS = {"AA": -22.2, "TT": -22.2, "AT": -20.4, "TA": -21.3, "CA": -22.7, "TG": -22.7, "GT": -22.4, "AC": -22.4,
       "CT": -21.0, "AG": -21.0, "GA": -22.2, "TC": -22.2, "CG": -27.2, "GC": -24.4, "GG": -19.9, "CC": -19.9}


def _terminal(s):
    """Initiation terms for the two ends (SantaLucia 1998): terminal G.C 0.1 / -2.8, terminal A.T 2.3 / 4.1."""
    dh = ds = 0.0
    for end in (s[0], s[-1]):
        if end in "GC":
            dh += 0.1
            ds += -2.8
        elif end in "AT":
            dh += 2.3
            ds += 4.1
    return dh, ds


def temp(s, DNA_c=5000.0, Na_c=10.0, Mg_c=20.0, dNTPs_c=10.0, uncorrected=False):
    """DNA/DNA melting temperature in deg C.  DNA_c in nM; Na_c, Mg_c, dNTPs_c in mM."""
    s = s.upper()
    if len(s) < 2 or any(c not in "ACGT" for c in s):
        raise ValueError("melting.temp needs an ACGT sequence of at least 2 bases: %r" % s)
    R = 1.987                                            # cal / (K mol)
    dh, ds = _terminal(s)
    for i in range(len(s) - 1):                          # every (overlapping) dinucleotide
        dh += _DH[s[i:i + 2]]
        ds += _DS[s[i:i + 2]]
    k = DNA_c * 1e-9
    tm = (1000.0 * dh) / (ds + R * log(k))               # kelvin
    if uncorrected:
        return tm - 273.15
    fgc = sum(1 for c in s if c in "GC") / float(len(s))
    m_na, m_mg, m_dntp = Na_c * 1e-3, Mg_c * 1e-3, dNTPs_c * 1e-3
    # free Mg2+ (dNTPs bind it), Owczarzy et al. (2008) eq. 16
    ka = 3e4
    d = (ka * m_dntp - ka * m_mg + 1.0) ** 2 + 4.0 * ka * m_mg
    f_mg = (-(ka * m_dntp - ka * m_mg + 1.0) + sqrt(d)) / (2.0 * ka)
    ratio = sqrt(f_mg) / m_na if m_na > 0 else 7.0
    if ratio < 0.22:                                     # monovalent ions dominate: Owczarzy (2004) eq. 4
        ln = log(m_na)
        inv = 1.0 / tm + ((4.29 * fgc - 3.95) * ln + 0.940 * ln * ln) * 1e-5
    else:
        a, dd, g = 3.92, 1.42, 8.31
        if ratio < 6.0:                                  # both compete: Na-dependent coefficients (eqs. 18-20)
            ln = log(m_na)
            a *= 0.843 - 0.352 * sqrt(m_na) * ln
            dd *= 1.279 - 4.03e-3 * ln - 8.03e-3 * ln * ln
            g *= 0.486 - 0.258 * ln + 5.25e-3 * ln ** 3
        lm = log(f_mg)
        inv = 1.0 / tm + (a - 0.911 * lm + fgc * (6.26 + dd * lm)
                          + (1.0 / (2.0 * (len(s) - 1))) * (-48.2 + 52.5 * lm + g * lm * lm)) * 1e-5
    return 1.0 / inv - 273.15
