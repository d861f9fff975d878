#!/usr/bin/env python
"""bench.py -- NRMS hot-path benchmark on B200 (contract: see task brief / DESIGN.md section 6).

A "step" is ONE full evaluate pass (BASELINE.json configs[1]) over synthetic MIND-small-shaped data
per GPU: encode all 65,238 news, user vectors for 73,152 impressions, score every candidate, rank
metrics (AUC/MRR/nDCG@5/@10).  Weak scaling: every rank owns one MIND-small-shaped shard (N x news,
N x impressions in total); the news-vector table is all-gathered over NCCL, metric sums all-reduced.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision tf32|fp32] [--impl reference]

Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference's
path (oracle/, numpy on all host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

# torch reads this when its CUDA allocator starts: VMM-mapped (2 MB page) segments, see _lib._prefer_large_page_segments
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NEWS_PER_GPU = 65238
IMPRESSIONS_PER_GPU = 73152
NUM_WORDS = 70976
# algorithmic work per unit (SURVEY.md 8(d) / DESIGN.md section 5)
FLOP_PER_TITLE = 13_700_000
FLOP_PER_USER = 36_050_000
BYTES_PER_CANDIDATE = 1212          # fp32 news vector + int64 index + fp32 score (FP32 mode)
BYTES_PER_CANDIDATE_F16 = 652       # tensor mode: the candidate row is read from the fp16 table copy (640 B)
BYTES_PER_IMPRESSION = 1208
# K1g per user: 50 gathered rows x 1,800 B (q|k|v fp16, the useful bytes of the 2,160-byte padded row) + 50 int32
# history indices + 50 x 300 fp16 context values written for K2
K1G_BYTES_PER_USER = 50 * 1800 + 50 * 4 + 50 * 600


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "reasons": sorted(reasons), "samples": len(sm)}


MIND_LARGE = (161013, 376471)      # BASELINE configs[3]: news, impressions (strong scaling: fixed total)


def workload_sizes(world, workload="small-per-gpu"):
    """(news, impressions) in total over all ranks."""
    if workload == "mind-large":
        return MIND_LARGE
    return NEWS_PER_GPU * world, IMPRESSIONS_PER_GPU * world


def make_data(world, workload="small-per-gpu"):
    from newsrecommendationsystem_b200 import synthetic
    n_news, n_imp = workload_sizes(world, workload)
    news = synthetic.make_news(n_news, num_words=NUM_WORDS, seed=1234)
    imp = synthetic.make_impressions(n_imp, n_news, seed=1234)
    return news, imp


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path, timed on a bounded sample and extrapolated
# --------------------------------------------------------------------------------------------------
def cpu_baseline(news, imp, n_titles=8192, n_users=2048, n_score=6000):
    """The reference's CPU path (torch-CPU port of its op sequence, oracle/torch_port.py; all host
    cores) on a bounded sample of the workload, extrapolated to the full MIND-small-shaped pass:
    get_news_vector in 2,048-title batches (evaluate.py:187), get_user_vector in 2,048-user
    batches built by tensor indexing, get_prediction + .tolist() + calculate_single_user_metric
    per impression (evaluate.py:245-265, :160-168)."""
    import torch
    from newsrecommendationsystem_b200 import synthetic
    from oracle import nrms_oracle as O
    from oracle import torch_port as TP
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic.init_state_dict(num_words=NUM_WORDS, seed=0)
    TP.news_vectors(sd, news[:256])     # warm-up
    t0 = time.perf_counter()
    for s in range(0, n_titles, 2048):
        TP.news_vectors(sd, news[s:s + 2048])
    t_news = (time.perf_counter() - t0) / n_titles
    rng = np.random.default_rng(0)
    table = torch.from_numpy(rng.standard_normal((4097, 300)).astype(np.float32) * 0.3)
    table[4096] = 0
    hist = torch.from_numpy(np.where(imp["hist_rows"][:n_users] < 0, 4096, imp["hist_rows"][:n_users] % 4096))
    TP.user_vectors(sd, table[hist[:64]])
    t0 = time.perf_counter()
    uvs = []
    for s in range(0, n_users, 2048):
        uvs.append(TP.user_vectors(sd, table[hist[s:s + 2048]]))
    t_user = (time.perf_counter() - t0) / n_users
    uv = torch.cat(uvs)
    offs = imp["cand_offsets"]
    cand = torch.from_numpy(imp["cand_rows"][:int(offs[n_score])] % 4096)
    t0 = time.perf_counter()
    for i in range(n_score):
        a, b = int(offs[i]), int(offs[i + 1])
        y_pred = TP.prediction(table[cand[a:b]], uv[i % n_users])
        O.single_user_metric(imp["labels"][a:b], y_pred)
    t_score = (time.perf_counter() - t0) / n_score
    n_news, n_imp = NEWS_PER_GPU, IMPRESSIONS_PER_GPU
    total = n_news * t_news + n_imp * t_user + n_imp * t_score
    return dict(value=n_imp / total, unit="impressions/s", cores=cores, kind="port",
                sample=f"torch-CPU port of the reference op sequence: {n_titles} titles + {n_users} users + {n_score} "
                       f"impressions scored+ranked, extrapolated to {n_news} news / {n_imp} impressions",
                news_per_s=1.0 / t_news, users_per_s=1.0 / t_user, scored_impressions_per_s=1.0 / t_score)


def run_reference(args, rank):
    if rank != 0:
        return
    news, imp = make_data(1)
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        last = cpu_baseline(news, imp, n_titles=4096, n_users=2048, n_score=3000)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            vals.append((last["value"], dt))
    v = float(np.mean([x[0] for x in vals]))
    line = dict(impl="reference", metric="evaluate impressions/s", value=v, unit="impressions/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * IMPRESSIONS_PER_GPU / v,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=workload_config(args.gpus, "cpu"),
                cpu_baseline=dict(value=v, unit="impressions/s", cores=last["cores"], kind="port", sample=last["sample"]),
                e2e=dict(value=v, unit="impressions/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(world, precision, workload="small-per-gpu"):
    n_news, n_imp = workload_sizes(world, workload)
    name = ("NRMS evaluate pipeline, MIND-large-shaped in total, sharded over the GPUs (BASELINE configs[3])"
            if workload == "mind-large" else "NRMS evaluate pipeline, MIND-small-shaped per GPU (BASELINE configs[1])")
    return dict(workload=name, news_per_gpu=n_news / world, impressions_per_gpu=n_imp / world, vocab=NUM_WORDS,
                title_len=20, history=50, heads=15, dim=300, precision=precision,
                parallelism=f"dp{world}: news rows + impressions sharded, NCCL all-gather of the news table",
                l2="flushed between timed steps (256 MiB write); per-step CUDA events summed")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--workload", default="small-per-gpu", choices=["small-per-gpu", "mind-large"],
                    help="small-per-gpu (default, weak scaling: one MIND-small-shaped shard per GPU, the headline) or "
                         "mind-large (strong scaling: 161,013 news / 376,471 impressions in total, BASELINE configs[3])")
    ap.add_argument("--no-train", action="store_true", help="skip the train-step side measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from newsrecommendationsystem_b200 import NRMS, NRMSConfig, synthetic, _lib
    from newsrecommendationsystem_b200.evaluate import EvalHost, EvalInputs, evaluate_tensors
    from newsrecommendationsystem_b200.train import TrainStep

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # ---- model + data (identical on every rank) ------------------------------------------------
    sd = synthetic.init_state_dict(num_words=NUM_WORDS, seed=0)
    model = NRMS(NRMSConfig)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model.to(dev).eval().set_precision(args.precision)
    news, imp = make_data(world, args.workload)
    news_total, _ = workload_sizes(world, args.workload)
    host = EvalHost(news, imp["hist_rows"], imp["cand_offsets"], imp["cand_rows"], imp["labels"])
    inputs = EvalInputs.from_host(host, dev)
    n_imp_total = host.n_impressions
    n_cand_total = int(host.cand_offsets_host[-1])
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stage_ms = {}
    h2d_bytes = [host.nbytes()]

    def timed_eval(resident, record_stages=False):
        marks = []

        def mark(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))
        flush_buf.fill_(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        inp = resident if resident is not None else EvalInputs.from_host(host, dev)
        h2d_bytes[0] = inp.h2d_bytes
        means = evaluate_tensors(model, inp, mark=mark if record_stages else None)   # ends with the D2H of 8 doubles
        e1.record()
        torch.cuda.synchronize()
        if record_stages:
            for (n0, a), (n1, b) in zip(marks[:-1], marks[1:]):
                stage_ms.setdefault(n1, []).append(a.elapsed_time(b))
        return e0.elapsed_time(e1), means

    for w in range(args.warmup):
        if w == args.warmup - 1:
            lib.nrms_set_option(b"time_k1", 1)     # the timing events are created (and pooled) outside the timed region
        timed_eval(inputs)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    lib.nrms_set_option(b"time_k1", 1)     # CUDA events around every user-encoder K1 launch of the timed region
    l0 = lib.nrms_launch_count()
    times = []
    means = None
    for _ in range(args.steps):
        ms, means = timed_eval(inputs, record_stages=True)
        times.append(ms)
    launches = int(lib.nrms_launch_count() - l0)
    kstat = {k: tuple(lib.nrms_get_stat(f"{k}_{w}".encode()) for w in ("ms", "launches", "sequences"))
             for k in ("k1", "k1n", "k1g", "k1gn")}
    lib.nrms_set_option(b"time_k1", 0)
    barrier()
    e2e_times = []
    for i in range(args.steps + 1):
        ms, _ = timed_eval(None)
        if i > 0:
            e2e_times.append(ms)
    clocks = sampler.stop() if sampler else None

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total_ms = max_over_ranks(float(np.sum(times)))
    e2e_ms = max_over_ranks(float(np.sum(e2e_times)))
    ms_per_step = total_ms / args.steps
    value = n_imp_total * args.steps / (total_ms / 1e3)
    e2e_value = n_imp_total * args.steps / (e2e_ms / 1e3)
    st = {k: float(np.mean(v)) for k, v in stage_ms.items()}

    # ---- roofline of the dominant stage --------------------------------------------------------
    peaks = load_peaks()
    n_news_rank = news_total / world       # per-rank shard
    n_imp_rank = n_imp_total / world
    cand_rank = n_cand_total / world
    stage_flops = {"news": n_news_rank * FLOP_PER_TITLE, "users": n_imp_rank * FLOP_PER_USER}
    # Dominant kernel = the user-encoder attention kernel.  Its average launch duration is measured live (CUDA events
    # on the launching stream, nrms_set_option("time_k1")).  Tensor mode runs K1g (k1g_table_attn.cu): the q|k|v rows of
    # a pre-projected table are gathered per history row, so the kernel is bound by bytes moved, not by the projection
    # GEMM it no longer contains; FP32-table runs (user_table_attn = 0) keep K1 v6, bound by the tensor pipe.
    tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))) \
        if os.path.exists(os.path.join(ROOT, "profiles", "r1_traffic.json")) else {}

    def tensor_roof(kind, name, flop_per_seq, traffic_key):
        ms, n, seqs = kstat[kind]
        if not (n > 0 and ms > 0):
            return None
        us = 1e3 * ms / n
        ach = (seqs / n) * flop_per_seq / (us * 1e-6) / 1e12
        return dict(bound="tensor", kernel=name, achieved=ach, peak=peaks["bf16_tflops"], unit="TFLOP/s",
                    frac=ach / peaks["bf16_tflops"], traffic=tr.get(traffic_key), us_per_launch=us, launches=int(n),
                    sequences_per_launch=seqs / n, algorithmic_flop_per_sequence=flop_per_seq,
                    peak_source=f"{peaks['source']} dense bf16/fp16 burst (operands are fp16, fp32 accumulate)")

    # K1 owns everything of an encoder except the additive projection/pooling (K2)
    roof_news = tensor_roof("k1n", "k1v6::encoder_attn_tc6_kernel<20,24,5> (news encoder: gather+QKV+attention)",
                            FLOP_PER_TITLE - 2_400_000 - 20_000, "encoder_attn_tc6_kernel<20,24,5>")
    def gather_roof(kind, name, bytes_per_seq, traffic_key):
        ms, n, seqs = kstat[kind]
        if not (n > 0 and ms > 0):
            return None
        us = 1e3 * ms / n
        ach = (seqs / n) * bytes_per_seq / (us * 1e-6) / 1e9
        return dict(bound="hbm", kernel=name, achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s", frac=ach / peaks["hbm_gbs"],
                    traffic=tr.get(traffic_key), us_per_launch=us, launches=int(n), sequences_per_launch=seqs / n,
                    algorithmic_bytes_per_sequence=bytes_per_seq,
                    peak_source=f"{peaks['source']} HBM copy bandwidth (burst); the projected table is partly "
                                f"L2-resident, see traffic")

    roof = gather_roof("k1g", "k1g::seq_attn_kernel<50> (user encoder: q|k|v row gather + 15-head attention)",
                       K1G_BYTES_PER_USER, "seq_attn_kernel<50>")
    if roof_news is None:       # the news encoder took the table path too: 20 token rows of the projected embedding table
        roof_news = gather_roof("k1gn", "k1g::seq_attn_kernel<20> (news encoder: q|k|v row gather + 15-head attention)",
                                20 * 1800 + 20 * 8 + 20 * 600, "seq_attn_kernel<20>")
    if roof is None:
        roof = tensor_roof("k1", "k1v6::encoder_attn_tc6_kernel<50,64,2> (user encoder: gather+QKV+attention)",
                           FLOP_PER_USER - 6_050_000, "encoder_attn_tc6_kernel<50,64,2>")
    if roof is None:
        dominant = max(st, key=st.get) if st else "news"
        ach = stage_flops.get(dominant, 0.0) / (st[dominant] / 1e3) / 1e12
        roof = dict(bound="tensor", kernel=f"{dominant} encoder stage", achieved=ach, peak=peaks["bf16_tflops"],
                    unit="TFLOP/s", frac=ach / peaks["bf16_tflops"], traffic=None, peak_source=peaks["source"])
    per_cand = BYTES_PER_CANDIDATE_F16 if args.precision == "tf32" else BYTES_PER_CANDIDATE
    # + the one-off fp16 copy of the table inside the stage (read fp32, write fp16)
    pack_bytes = (news_total + 1) * (1200 + 640) if args.precision == "tf32" else 0
    score_bytes = cand_rank * per_cand + n_imp_rank * BYTES_PER_IMPRESSION + pack_bytes
    extras = dict(
        stage_ms=st,
        news_per_s=n_news_rank * world / (st.get("news", float("nan")) / 1e3),
        users_per_s=n_imp_total / (st.get("users", float("nan")) / 1e3),
        score_candidates_per_s=n_cand_total / (st.get("score", float("nan")) / 1e3),
        score_hbm_gbs=score_bytes / (st.get("score", float("nan")) / 1e3) / 1e9,
        score_hbm_frac=score_bytes / (st.get("score", float("nan")) / 1e3) / 1e9 / peaks["hbm_gbs"],
        news_tflops=stage_flops["news"] / (st.get("news", float("nan")) / 1e3) / 1e12,
        users_tflops=stage_flops["users"] / (st.get("users", float("nan")) / 1e3) / 1e12,
        metrics=dict(zip(("auc", "mrr", "ndcg5", "ndcg10"), means)),
        roofline_news=roof_news,
    )

    # ---- training step side measurement (BASELINE configs[2]: B=128, 1+4 candidates) ------------
    train = None
    if not args.no_train:
        model.train()
        ts = TrainStep(model, lr=1e-4)
        cand, clicked = synthetic.make_train_batch(128, news[:NEWS_PER_GPU], k_neg=4, seed=1234 + rank)
        titles = torch.from_numpy(np.concatenate([cand, clicked], axis=1)).pin_memory()
        for _ in range(3):
            ts.step_tokens(titles, 5)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_train = max(5, args.steps)
        e0.record()
        for _ in range(n_train):
            loss = ts.step_tokens(titles, 5)
        loss_val = float(loss.item())            # D2H of the loss, like train.py:225
        e1.record()
        torch.cuda.synchronize()
        tms = max_over_ranks(e0.elapsed_time(e1))
        train = dict(samples_per_s=128 * world * n_train / (tms / 1e3), ms_per_step=tms / n_train, batch_per_gpu=128,
                     k_neg=4, dropout=0.2, loss=loss_val, forward_precision=args.precision,
                     backward_precision=args.precision)
        model.eval()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(news, imp)

    if rank == 0:
        line = dict(metric="evaluate impressions/s", value=value, unit="impressions/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms_per_step, higher_is_better=True,
                    scaling="strong" if args.workload == "mind-large" else "weak",
                    vs_baseline=None, dtype="f16xf16->f32 (tcgen05 kind::f16; TF32-equivalent 11-bit significand)"
                    if args.precision == "tf32" else "f32", data="synthetic",
                    config=workload_config(world, args.precision, args.workload), clocks=clocks,
                    e2e=dict(value=e2e_value, unit="impressions/s", h2d_bytes_per_step=int(h2d_bytes[0]),
                             d2h_bytes_per_step=64, ms_per_step=e2e_ms / args.steps),
                    gpu_launches=launches, roofline=roof, cpu_baseline=cpu, train=train, **extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
