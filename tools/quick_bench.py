"""Developer sweep: kernel-only GCUPS of the PairHMM fast path for several K (rows per lane)."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import lorikeet_b200 as lk
    from lorikeet_b200 import synth
    n = int(sys.argv[2])
    b = synth.pairhmm_regions(n)
    with lk.Context(0) as ctx:
        res = ctx.pairhmm_upload(b)
        for _ in range(2):
            res.run()
        ms = [res.run() for _ in range(5)]
        out, flags = res.download(want_flags=True)
        best = min(ms)
        import hashlib
        print("K=%s regions=%d cells=%.3e best_ms=%.3f GCUPS=%.1f launches=%d rescued=%d sha=%s" % (
            os.environ.get("LK_FORCE_K", "auto"), n, res.n_cells, best, res.n_cells / best / 1e6, res.launch_count,
            int(flags.sum()), hashlib.sha1(out.tobytes()).hexdigest()[:12]), flush=True)
        t = time.time()
        ctx.pairhmm_batch(b)
        print("   e2e host call: %.1f ms" % ((time.time() - t) * 1e3), flush=True)
else:
    n = sys.argv[1] if len(sys.argv) > 1 else "2000"
    for k in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["p0", "p5", "p6", "p8", "p10", "s0", "s19"]):
        env = dict(os.environ)
        k = k.lstrip("ps")
        if k != "0":
            env["LK_FORCE_K"] = k
        subprocess.run([sys.executable, __file__, "child", n], env=env)
