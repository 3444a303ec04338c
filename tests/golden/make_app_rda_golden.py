#!/usr/bin/env python
"""Generate tests/golden/app_rda_golden.npz from the reference's shipped fitted objects.

Run in the build container only (it reads /root/reference/data/app.rda, which does not exist on the GPU
box):  python tests/golden/make_app_rda_golden.py

data/app.rda (xz, R serialisation "RDX2", XDR big-endian) holds app = list of 3 data sets x k=1..5 fitted
`mnem` objects (other/mnem_data_raw.r:311-318).  For each object that still carries `probs` (k <= 3) we
keep (probs k x C, mw, ll): ll = getLL(probs, mw) is a known answer for getLL / getAffinity
(R/mnems_low.r:540-552, R/mnems.r:1390-1441).  For every object we keep comp[[i]]$phi / $theta as
realistic graph inputs for the closure / reduction tests.
"""
import lzma
import struct
import sys

import numpy as np

SRC = "/root/reference/data/app.rda"


class Reader:
    def __init__(self, buf):
        self.b, self.p, self.refs = buf, 0, []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def f64(self, n):
        a = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
        self.p += 8 * n
        return a

    def i32s(self, n):
        a = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
        self.p += 4 * n
        return a

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & (1 << 9)), bool(flags & (1 << 10))
        if t == 254:      # NILVALUE
            return None
        if t == 255:      # REFSXP
            return self.refs[(flags >> 8) - 1]
        if t == 1:        # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:        # LISTSXP (pairlist): returns list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None  # noqa: F841
                tag = self.item() if has_tag else None
                out.append((tag, self.item()))
                flags = self.i32()
                t = flags & 0xFF
                if t == 254:
                    return out
                assert t == 2, t
                has_attr, has_tag = bool(flags & (1 << 9)), bool(flags & (1 << 10))
        if t == 9:        # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.p:self.p + n].decode("latin1")
            self.p += n
            return s
        if t in (10, 13):
            val = self.i32s(self.i32())
        elif t == 14:
            val = self.f64(self.i32())
        elif t == 16:
            val = [self.item() for _ in range(self.i32())]
        elif t == 19:
            val = [self.item() for _ in range(self.i32())]
        else:
            raise ValueError(f"unhandled SEXP type {t} at {self.p}")
        attrs = dict(self.item()) if has_attr else {}
        if t == 19 and "names" in attrs:
            return {"__names__": attrs["names"], **{n: v for n, v in zip(attrs["names"], val)}}
        if isinstance(val, np.ndarray) and "dim" in attrs:
            val = val.reshape(tuple(int(d) for d in attrs["dim"]), order="F")
        return val


def main():
    raw = lzma.open(SRC).read()
    assert raw[:5] == b"RDX2\n" and raw[5:7] == b"X\n", raw[:8]
    r = Reader(raw)
    r.p = 7
    r.i32(), r.i32(), r.i32()
    top = dict(r.item())
    app = top["app"]
    names = ["cropseq", "perturbseq_cc7d", "perturbseq_p7d"]   # order of other/mnem_data_raw.r:311-318
    assert len(app) == 3
    out = {}
    n_ll = 0
    for dname, fits in zip(names, app):
        for ki, fit in enumerate(fits, start=1):
            key = f"{dname}_k{ki}"
            for ci, c in enumerate(fit["comp"]):
                out[f"{key}_phi{ci}"] = np.asarray(c["phi"], dtype=np.float64)
                out[f"{key}_theta{ci}"] = np.asarray(c["theta"], dtype=np.int32)
            out[f"{key}_ll"] = np.asarray(fit["ll"], dtype=np.float64).reshape(-1)
            out[f"{key}_mw"] = np.asarray(fit["mw"], dtype=np.float64).reshape(-1)
            probs = fit.get("probs")
            if isinstance(probs, np.ndarray) and probs.ndim == 2 and probs.shape[0] == ki:
                out[f"{key}_probs"] = probs
                n_ll += 1
    out["__datasets__"] = np.array(names)
    np.savez_compressed("tests/golden/app_rda_golden.npz", **out)
    print(f"datasets={names} objects_with_probs={n_ll} arrays={len(out)}", file=sys.stderr)


if __name__ == "__main__":
    main()
