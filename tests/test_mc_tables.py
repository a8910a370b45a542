"""CPU: the marching-cubes case tables are derived (scripts/gen_mc_tables.py), not transcribed; against the reference's own tables
(parsed from /root/reference when present) they have, in all 256 cases, the same crossed edges, the same number of triangles and
the same oriented patch boundaries -- only the interior diagonals of patches with more than three vertices differ."""
import collections
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
REF = "/root/reference/src/core/cuda/TSDF.cu"


def test_tables_are_consistent():
    import gen_mc_tables as g
    et, tt = g.generate()
    assert et[0] == 0 and et[255] == 0 and tt[0] == [] and tt[255] == []
    for cfg in range(256):
        used = {e for t in tt[cfg] for e in t}
        assert used == {e for e in range(12) if (et[cfg] >> e) & 1}, cfg          # every crossed edge carries a vertex, no other
        assert et[cfg] == et[255 - cfg]                                             # complement: same edges
    hdr = open(os.path.join(ROOT, "emfusion_b200", "csrc", "mc_tables.h")).read()
    assert f"0x{et[1]:03x}" in hdr and "GENERATED" in hdr


@pytest.mark.skipif(not os.path.exists(REF), reason="reference not present on this box")
def test_tables_against_reference():
    import gen_mc_tables as g
    src = open(REF).read()
    m = re.search(r"edgeTable\s*\[\s*256\s*\]\s*=\s*\{(.*?)\};", src, re.S)
    et_r = [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(1))]
    m = re.search(r"triTable\s*\[\s*256\s*\]\s*\[\s*16\s*\]\s*=\s*\{(.*?)\};", src, re.S)
    tt_r = []
    for r in re.findall(r"\{([^{}]*)\}", m.group(1)):
        v = [int(x) for x in re.findall(r"-?\d+", r)]
        v = v[:v.index(-1)] if -1 in v else v
        tt_r.append([tuple(v[i:i + 3]) for i in range(0, len(v), 3)])
    et, tt = g.generate()

    def boundary(tris):      # directed edges that belong to exactly one triangle
        d, u = collections.Counter(), collections.Counter()
        for t in tris:
            for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                d[(a, b)] += 1; u[frozenset((a, b))] += 1
        return sorted(k for k in d if u[frozenset(k)] == 1)

    assert et == et_r
    assert [len(t) for t in tt] == [len(t) for t in tt_r]
    assert all(boundary(a) == boundary(b) for a, b in zip(tt, tt_r))
