import os, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic, _lib
from newsrecommendationsystem_b200.evaluate import EvalHost, EvalInputs, evaluate_tensors
lib = _lib.load()
dev = torch.device("cuda", 0)
sd = synthetic.init_state_dict(num_words=bench.NUM_WORDS, seed=0)
model = NRMS(NRMSConfig); model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); model.to(dev).eval().set_precision("tf32")
news, imp = bench.make_data(1)
host = EvalHost(news, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"])
inputs = EvalInputs.from_host(host, dev)
_mb = int(os.environ.get("PROBE_FLUSHBUF_MB", "256"))
flush = torch.empty(_mb << 20, dtype=torch.uint8, device=dev) if _mb > 0 and os.environ.get("PROBE_NOFLUSHBUF", "0") == "0" else None
for flag in (0, 1, 2, 3):      # 0 = K1 v6 everywhere, 1 = K1g (first S=50 kernel), 2 = templated K1g users, 3 = + news table path
    lib.nrms_set_option(b"user_table_attn", 1 if flag else 0)
    lib.nrms_set_option(b"k1g_variant", 2 if flag >= 2 else 0)
    lib.nrms_set_option(b"news_table_attn", 1 if flag == 3 else 0)
    for rep in range(3):
        if os.environ.get("PROBE_SYNC", "0") != "0": torch.cuda.synchronize()
        if os.environ.get("PROBE_FLUSH", "0") != "0" and flush is not None: flush.fill_(1)
        if rep >= 1 and os.environ.get("PROBE_TIME_K1", "1") != "0": lib.nrms_set_option(b"time_k1", 1)      # rep 1 creates the pooled events, rep 2 is the one read
        ev = {}
        host_t = {}
        def mark(n): e = torch.cuda.Event(enable_timing=True); e.record(); ev[n] = e; host_t[n] = time.perf_counter()
        means, det = evaluate_tensors(model, inputs, return_details=True, mark=mark)
    torch.cuda.synchronize()
    k1g_ms, k1_ms = lib.nrms_get_stat(b"k1g_ms"), lib.nrms_get_stat(b"k1_ms")
    print(f"  news attention kernel total: {lib.nrms_get_stat(b'k1n_ms') + lib.nrms_get_stat(b'k1gn_ms'):.3f} ms")
    lib.nrms_set_option(b"time_k1", 0)
    print(f"  user attention kernel total: {k1g_ms + k1_ms:.3f} ms")
    names = list(ev)
    print("  host ms between marks (no sync):", {names[i + 1]: round(1e3 * (host_t[names[i + 1]] - host_t[names[i]]), 3) for i in range(len(names) - 1)})
    print("table_attn", flag, {names[i + 1]: round(ev[names[i]].elapsed_time(ev[names[i + 1]]), 3) for i in range(len(names) - 1)}, means)
    uv = det["user_vectors"].clone()
    if flag == 0: uv0 = uv
    else:
        d = (uv - uv0).double().norm(dim=1) / uv0.double().norm(dim=1)
        print("  max rel diff vs the per-user path", float(d.max()))
