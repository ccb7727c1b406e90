import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np
import kmergma_jl_b200 as K
path = "/tmp/ingest_bench.fasta"
n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 800_000_000
if not os.path.exists(path) or os.path.getsize(path) < n:
    rng = np.random.default_rng(1)
    with open(path, "wb") as fh:
        for r in range(4):
            L = n // 4
            seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=L)].copy()
            seq[:10000] = ord("N"); seq[L // 2:L // 2 + 3_000_000] = ord("N"); seq[-10000:] = ord("N")
            if r == 1: seq[100000:200000] |= 0x20
            body = seq[:L // 80 * 80].reshape(-1, 80)
            fh.write(b">contig%d some description\n" % r)
            fh.write(np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1).tobytes())
            fh.write(seq[body.size:].tobytes() + b"\n")
for i in range(4):
    t0 = time.perf_counter()
    g = K.Genome.from_fasta(path)
    t1 = time.perf_counter()
    print("from_fasta %.1f ms  %.2f GB/s  total_len %d" % ((t1 - t0) * 1e3, os.path.getsize(path) / (t1 - t0) / 1e9, g.total_len))
    if i == 0:
        import zlib
        h = 0
        for r in range(len(g)):
            L = g.seqsize(r)
            h = zlib.crc32(g.seq(r, 1, min(L, 5_000_000)).encode(), h); h = zlib.crc32(g.seq(r, max(1, L - 5_000_000), L).encode(), h)
            h = zlib.crc32(g.seq(r, L // 2 - 100, L // 2 + 3_000_100).encode(), h)
        print("crc", h, "runs", g.masked_runs().shape, g.masked_runs()[:3].tolist())
    del g
