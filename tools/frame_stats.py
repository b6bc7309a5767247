"""Parse a Zstandard frame far enough to report, per block: type, literals (regenerated, stored bytes,
mode) and sequence count / section bytes.  Tuning aid for the encoder (test infrastructure)."""
def stats(fr):
    p = 4
    fhd = fr[p]; p += 1
    ss = (fhd >> 5) & 1
    if not ss: p += 1
    did = fhd & 3
    p += [0, 1, 2, 4][did]
    fcs = fhd >> 6
    p += [1 if ss else 0, 2, 4, 8][fcs]
    out = []
    while True:
        h = fr[p] | fr[p + 1] << 8 | fr[p + 2] << 16; p += 3
        last, bt, bs = h & 1, (h >> 1) & 3, h >> 3
        if bt == 0: out.append(dict(type="raw", size=bs)); p += bs
        elif bt == 1: out.append(dict(type="rle", size=bs)); p += 1
        else:
            b = fr[p:p + bs]
            lt = b[0] & 3; sf = (b[0] >> 2) & 3
            if lt < 2:
                if sf in (0, 2): regen = b[0] >> 3; hs = 1
                elif sf == 1: regen = (b[0] >> 4) | b[1] << 4; hs = 2
                else: regen = (b[0] >> 4) | b[1] << 4 | b[2] << 12; hs = 3
                comp = regen if lt == 0 else 1
            else:
                if sf < 2: v = b[0] | b[1] << 8 | b[2] << 16; regen = (v >> 4) & 1023; comp = v >> 14; hs = 3
                elif sf == 2: v = int.from_bytes(b[:4], "little"); regen = (v >> 4) & 16383; comp = v >> 18; hs = 4
                else: v = int.from_bytes(b[:5], "little"); regen = (v >> 4) & 262143; comp = v >> 22; hs = 5
            q = hs + comp
            n0 = b[q]
            if n0 < 128: nseq = n0; q += 1
            elif n0 < 255: nseq = ((n0 - 128) << 8) + b[q + 1]; q += 2
            else: nseq = b[q + 1] + (b[q + 2] << 8) + 0x7F00; q += 3
            modes = b[q] if nseq else 0
            out.append(dict(type="cmp", size=bs, lit_mode=lt, lit_regen=regen, lit_bytes=hs + comp, nseq=nseq, seq_bytes=bs - hs - comp, modes=modes))
            p += bs
        if last: break
    return out
