#!/usr/bin/env python3
"""Localise a scan mismatch: survivors vs numpy predicate, orderings vs numpy stable argsort."""
import gzip, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kubevirt-gpu-device-plugin_b200"))
import numpy as np
import kvgpu
from oracle import oracle as O
text = gzip.open(os.path.join(ROOT, "tests", "golden", "pci.ids.gz"), "rb").read()
ids = O.nv_ids(text)
ctx = kvgpu.Context(0); ctx.pciids_load(text)
names = ctx.name_table(0, 65536)
def alive(r):
    return (r["vendor"] == 0x10de) & ((r["flags"] & 15) == 0) & ((r["driver"] == 1) | (r["driver"] == 2))
for n in [int(a) for a in sys.argv[1:]] or [100_000, 454_656, 454_657, 460_000, 909_312, 1_000_000]:
    recs = O.gen_pci(0, n, ids, 19)
    for rep in range(2):
        res = ctx.scan_pci(recs)
        a = alive(recs); exp = recs[a]
        S = len(res.survivors)
        msg = ["n=%d rep=%d S=%d exp=%d" % (n, rep, S, len(exp))]
        m = min(S, len(exp))
        bad = np.nonzero(res.survivors["addr"][:m] != exp["addr"][:m])[0]
        if len(bad): msg.append("first addr mismatch at surv %d (rec %d vs %d), nbad=%d" % (bad[0], res.survivors["addr"][bad[0]], exp["addr"][bad[0]], len(bad)))
        slot_bad = [i for i in range(0, m, max(1, m // 2000)) if res.name_at(int(res.survivors["name_slot"][i])) != names[int(res.survivors["device"][i])]]
        if slot_bad: msg.append("name_slot mismatches (sampled): %d first %d" % (len(slot_bad), slot_bad[0]))
        if S == len(exp) and not len(bad):
            for nm, keys, off, perm, field in (("dev", res.dev_keys, res.dev_off, res.dev_perm, "device"), ("grp", res.grp_keys, res.grp_off, res.grp_perm, "iommu_group")):
                want = np.argsort(res.survivors[field], kind="stable").astype(np.uint32)
                if not np.array_equal(perm, want): msg.append("%s perm differs at %d" % (nm, np.nonzero(perm != want)[0][0]))
                uk, cnt = np.unique(res.survivors[field], return_counts=True)
                if not np.array_equal(uk, keys): msg.append("%s keys differ (%d vs %d)" % (nm, len(keys), len(uk)))
                elif not np.array_equal(np.diff(off), cnt): msg.append("%s offsets differ" % nm)
        print("; ".join(msg))
