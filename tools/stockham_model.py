"""Model of the register-resident Stockham FFT used by the CUDA kernel.

Checks (1) the thread/register index mapping against numpy's FFT and (2) the
shared-memory bank-conflict degree of each exchange for a given padding rule.
Design aid only (not a test, not product code).
"""
import sys
import numpy as np

def plan(n2, e):
    radices = []
    rem = n2
    while rem > 1:
        r = min(e, rem)
        # prefer leaving a tail that is still >= 2
        radices.append(r)
        rem //= r
    return radices

def wavefronts(word_addrs_per_lane, bytes_per_lane):
    """Count smem wavefronts for one warp access: lanes -> list of 4B-word addresses."""
    # hardware processes 128B per wavefront; lanes hitting distinct banks go together
    lanes_per_phase = 128 // bytes_per_lane
    total = 0
    lanes = list(word_addrs_per_lane)
    for p in range(0, len(lanes), lanes_per_phase):
        grp = lanes[p:p + lanes_per_phase]
        bank_to_addrs = {}
        for words in grp:
            for w in words:
                bank_to_addrs.setdefault(w % 32, set()).add(w)
        total += max(len(s) for s in bank_to_addrs.values())
    return total

def simulate(n2, e, radices, pad, verbose=True):
    nt = n2 // e
    rng = np.random.default_rng(1)
    z = rng.standard_normal(n2) + 1j * rng.standard_normal(n2)
    # registers: v[t][m] <-> index t + m*nt
    v = np.array([[z[t + m * nt] for m in range(e)] for t in range(nt)])
    ns = 1
    report = []
    for pi, r in enumerate(radices):
        bpt = e // r
        out = np.zeros(n2, complex)
        widx = np.zeros((nt, e), int)
        for t in range(nt):
            for b in range(bpt):
                j = t + b * nt
                k = j % ns
                x = np.array([v[t][b + rr * bpt] * np.exp(-2j * np.pi * rr * k / (ns * r)) for rr in range(r)])
                y = np.fft.fft(x)
                j0 = (j // ns) * ns * r + k
                for rr in range(r):
                    out[j0 + rr * ns] = y[rr]
                    widx[t, b + rr * bpt] = j0 + rr * ns
        ns *= r
        last = pi == len(radices) - 1
        if last:
            # outputs must already be in place: v[m] <-> t + m*nt
            for t in range(nt):
                for m in range(e):
                    assert widx[t, m] == t + m * nt, (t, m, widx[t, m])
        # bank conflicts of the exchange (float2 = 2 words)
        if not last:
            wf_w = 0
            wf_r = 0
            for w0 in range(0, nt, 32):
                for m in range(e):
                    lanes = [[2 * pad(widx[t, m]), 2 * pad(widx[t, m]) + 1] for t in range(w0, min(w0 + 32, nt))]
                    wf_w += wavefronts(lanes, 8)
                    lanes = [[2 * pad(t + m * nt), 2 * pad(t + m * nt) + 1] for t in range(w0, min(w0 + 32, nt))]
                    wf_r += wavefronts(lanes, 8)
            ideal = (nt * e * 8 + 127) // 128
            report.append((pi, r, wf_w, wf_r, ideal))
        v = np.array([[out[t + m * nt] for m in range(e)] for t in range(nt)])
    ref = np.fft.fft(z)
    res = np.zeros(n2, complex)
    for t in range(nt):
        for m in range(e):
            res[t + m * nt] = v[t][m]
    err = np.abs(res - ref).max() / np.abs(ref).max()
    if verbose:
        print(f"N2={n2} E={e} NT={nt} radices={radices} err={err:.2e}")
        for pi, r, ww, wr, ideal in report:
            print(f"   exchange after pass {pi} (radix {r}): write wavefronts {ww}, read {wr}, ideal {ideal}")
    return err

if __name__ == "__main__":
    pads = {
        "none": lambda i: i,
        "i+i/16": lambda i: i + (i >> 4),
        "i+i/32": lambda i: i + (i >> 5),
        "i+i/8": lambda i: i + (i >> 3),
        "i+i/4": lambda i: i + (i >> 2),
    }
    cfgs = [(2048, 16, [16, 16, 8]), (1024, 16, [16, 16, 4]), (1024, 8, [8, 8, 4, 4]), (512, 8, [8, 8, 8]),
            (256, 8, [8, 8, 4]), (128, 4, [4, 4, 4, 2]), (2048, 8, [8, 8, 8, 4]), (512, 16, [16, 16, 2])]
    for n2, e, rad in cfgs:
        for name, pad in pads.items():
            print("pad", name, end=": ")
            simulate(n2, e, rad, pad)
