"""Is the user-encoder time data-dependent?  One GPU, the 2-rank weak workload (130 k news): the fp16 table of ALL news,
then the indexed user encoder over impression block 0, block 1, and block 1 with its users in reversed / shuffled order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic, _lib, ops
from newsrecommendationsystem_b200.evaluate import EvalHost, shard_impressions_by_candidates
lib = _lib.load()
dev = torch.device("cuda", 0)
sd = synthetic.init_state_dict(num_words=bench.NUM_WORDS, seed=0)
model = NRMS(NRMSConfig); model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); model.to(dev).eval().set_precision("tf32")
news, imp = bench.make_data(2)
host = EvalHost(news, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"])
with torch.no_grad():
    vec = model.news_encoder.encode_tokens(torch.from_numpy(news).to(dev))
    table = torch.cat([vec, torch.zeros(1, 300, device=dev)])
    table16 = ops.pack_rows_f16(table)
b = shard_impressions_by_candidates(host.cand_offsets_host, 2)
hist = host.hist_rows.to(dev)
lib.nrms_set_option(b"time_k1", 1)
def run(name, rows):
    for i in range(6):
        if i == 2:
            lib.nrms_set_option(b"time_k1", 1)
        model.user_encoder.forward_indexed(table16, rows)
    torch.cuda.synchronize()
    ms, n = lib.nrms_get_stat(b"k1g_ms"), lib.nrms_get_stat(b"k1g_launches")
    print(f"{name}: {1e3 * ms / n:.1f} us per K1g launch ({int(n)} launches, {rows.shape[0]} users, pads {(rows == host.n_news).float().mean().item():.3f})", flush=True)
blk0, blk1 = hist[b[0]:b[1]].contiguous(), hist[b[1]:b[2]].contiguous()
run("block 0", blk0)
run("block 1", blk1)
run("block 1 reversed", blk1.flip(0).contiguous())
run("block 0 shuffled", blk0[torch.randperm(blk0.shape[0], device=dev)].contiguous())
run("block 0 again", blk0)
print("hist row stats block0/1:", blk0.float().mean().item(), blk1.float().mean().item())

# ---- hot-row experiment: half of all gathered rows are PADDED_NEWS = ONE table row.  Replicate it R times (all zero
# vectors: same arithmetic) and spread the pad references over the replicas.
for R in (64, 1024):
    table_r = torch.cat([vec, torch.zeros(R, 300, device=dev)])
    table16_r = ops.pack_rows_f16(table_r)
    def spread(rows):
        rows = rows.clone()
        pad = rows == host.n_news
        idx = torch.arange(rows.numel(), device=dev, dtype=torch.int32).view_as(rows)
        rows[pad] = host.n_news + (idx[pad] * 7 + idx[pad] // 50) % R
        return rows
    table16 = table16_r
    u0 = model.user_encoder.forward_indexed(table16_r, spread(blk0))
    run(f"block 0, {R} pad replicas", spread(blk0))
    run(f"block 1, {R} pad replicas", spread(blk1))
table16 = ops.pack_rows_f16(table)
u_ref = model.user_encoder.forward_indexed(table16, blk0)
print("max |user vec (replicas) - user vec (one pad row)| =", float((u0 - u_ref).abs().max()))
