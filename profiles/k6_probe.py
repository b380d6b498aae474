"""K6 (map_variations) throughput probe: the (block, node) problems of one leaf merge -- 40 blocks of ~125 kbp, two node
sequences each at ~0.5 % from the consensus -- through pgmm_map_variations_batch, next to the oracle restatement of the
reference on the host cores.  usage: k6_probe.py [n_problems] [block_len] [cpu_sample]"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import naref  # noqa: E402
from test_nextalign_emul import mutate, rand_seq  # noqa: E402
from pangraph_b200 import abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
L = int(sys.argv[2]) if len(sys.argv) > 2 else 125_000
n_cpu = int(sys.argv[3]) if len(sys.argv) > 3 else 16
rng = np.random.default_rng(5)
refs, qrys = [], []
for k in range(n // 2):
    ref = rand_seq(rng, L + 1000 * (k % 7))
    for _ in range(2):
        refs.append(ref), qrys.append(mutate(rng, ref, sub=0.005, indel=0.0003, max_indel=5))
ms, bw = [0] * len(refs), [8] * len(refs)
abi.map_variations_batch(refs[:2], qrys[:2], ms[:2], bw[:2])  # warm-up (context, kernel image)
t0 = time.perf_counter()
got, st = abi.map_variations_batch(refs, qrys, ms, bw, with_stats=True)
wall = time.perf_counter() - t0
bp = sum(len(r) for r in refs)
print(f"GPU: {len(refs)} problems, {bp / 1e6:.1f} Mbp of consensus, {st['cells'] / 1e6:.0f} M band cells ({st['problems']:.0f} attempts, {st['launches']:.0f} launches): "
      f"kernel {st['kernel_ms']:.1f} ms = {st['cells'] / st['kernel_ms'] / 1e6:.2f} GCUPS, call {wall * 1e3:.1f} ms = {bp / wall / 1e9:.3f} Gbp/s")
sample = list(range(min(n_cpu, len(refs))))
t0 = time.perf_counter()
want0 = naref.map_variations(refs[0], qrys[0], 0, 8)
t1 = time.perf_counter() - t0
cores = len(os.sched_getaffinity(0))
t0 = time.perf_counter()
with ThreadPoolExecutor(cores) as ex:  # ctypes drops the GIL inside the C call
    want = list(ex.map(lambda i: naref.map_variations(refs[i], qrys[i], 0, 8), sample))
tall = time.perf_counter() - t0
cells1 = (len(refs[0]) + 1) * (2 * 13 + 1) * want0["attempts"]
print(f"CPU oracle (restatement of the reference): one problem on one core {t1 * 1e3:.1f} ms = {cells1 / t1 / 1e9:.3f} GCUPS; {len(sample)} problems on {cores} cores "
      f"{tall * 1e3:.1f} ms = {sum(len(refs[i]) for i in sample) / tall / 1e9:.3f} Gbp/s")
ok = all({k: got[i][k] for k in ('subs', 'dels', 'inss')} == {k: want[j][k] for k in ('subs', 'dels', 'inss')} for j, i in enumerate(sample))
print("parity on the CPU sample:", ok)
