#!/usr/bin/env python
"""Extract the E3SM hybrid vertical coordinate tables shipped with the reference into a small JSON data file.

Source (read-only reference tree, only readable in the build container):
  components/homme/test/vcoord/acme-72i.ascii, acme-72m.ascii  (L72: hyai,hybi / hyam,hybm)
  components/homme/test/vcoord/cami-30.ascii,  camm-30.ascii   (L30)
Output: acme-ecp_b200/data/vcoord.json  (committed; the GPU box has no /root/reference)
"""
import json, os, re, sys

REF = "/root/reference/components/homme/test/vcoord"
OUT = os.path.join(os.path.dirname(__file__), "..", "acme-ecp_b200", "data", "vcoord.json")


def read_ascii(path):
    toks = []
    with open(path) as f:
        for line in f:
            line = line.split("!")[0]
            toks += line.split()
    out, p = [], 0
    while p < len(toks):
        n = int(toks[p]); p += 1
        out.append([float(x) for x in toks[p:p + n]]); p += n
    return out


def main():
    res = {}
    for name, fi, fm in (("L72", "acme-72i.ascii", "acme-72m.ascii"), ("L30", "cami-30.ascii", "camm-30.ascii")):
        hyai, hybi = read_ascii(os.path.join(REF, fi))[:2]
        hyam, hybm = read_ascii(os.path.join(REF, fm))[:2]
        assert len(hyai) == len(hyam) + 1
        res[name] = dict(hyai=hyai, hybi=hybi, hyam=hyam, hybm=hybm)
    with open(OUT, "w") as f:
        json.dump(res, f)
    print("wrote", OUT, {k: len(v["hyam"]) for k, v in res.items()})


if __name__ == "__main__":
    main()
