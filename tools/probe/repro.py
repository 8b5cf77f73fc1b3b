import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]
import numpy as np, synth, krepp_b200, oracle_lib as O
toy = os.path.join(ROOT, "oracle", "_ref", "toy"); idx = os.path.join(toy, "index_toy")
seq, offs = synth.load_packed(os.path.join(toy, "genomes.npz"))
reads = synth.sample_reads(seq, offs, 20000, seed=1)
o = O.OracleIndex(idx)
s = reads[1405].tobytes()
want = [(m["strand"], m["leaf_se"], tuple(m["hist"])) for m in o.query(s)["minfo"]]
print("oracle", want)
for scan in ("lane", "staged"):
    os.environ["KREPP_SCAN"] = scan
    g = krepp_b200.Index(idx, 0)
    for n in (1, 3, 64):
        b = krepp_b200.IBatch(g, [s] * n)
        for rep in range(2):
            b.submit(); r = b.wait()
            got = [[(int(x["strand"]), int(x["leaf_se"]), tuple(int(v) for v in r["hist"][j])) for j, x in enumerate(r["records"]) if x["read"] == i] for i in range(n)]
            bad = [i for i in range(n) if got[i] != want]
            print(scan, "n", n, "rep", rep, "bad reads", bad[:8], "first got", got[0])
        b.close()
    g.close()
