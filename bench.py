#!/usr/bin/env python
"""bench.py — Mutations query throughput (seq·positions/s) on B200, with roofline and CPU baseline.

Workload (BASELINE.json configs[1], SURVEY.md §8d input 2): synthetic SARS-CoV-2-length table,
row i = evolved[i % |evolved|] (performance/sequence_generator.h tree model), 10 M rows x 29,903 nt
PER GPU (weak scaling; 8 GPUs = the 80 M-row config), query

    default.filter(date.between('2021-01-01','2021-06-30') && pango_lineage.lineage(<gen-2 node>,
                   includeSublineages:=true)).mutations(minProportion:=0.05, sequenceNames:={main})

A step = one such query: filter program (RangeSelection AND lineage IndexScan) -> dense filter ->
Mutations counts over every stored container of the touched chunks -> [NCCL allreduce of the
16 x 29,903 u32 counts] -> (e2e only) D2H + the reference's host-side thresholding.

  value  device-resident inputs (program prepared once), only kernels + collective in the timed region
  e2e    through the host layer / C ABI with HOST buffers: program H2D, cardinality + counts D2H,
         thresholding on the host, every step

`--impl reference` times the CPU restatement of the reference's path (oracle/, the reference itself
cannot be built in this image) on the box's host cores, one query per thread as the reference
deploys it (one Poco worker per request, no intra-query parallelism).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GENOME_LENGTH = 29903
N_SYMBOLS = 16
VALID_MUTATION_SYMBOLS = 5  # Nucleotide::VALID_MUTATION_SYMBOLS: -, A, C, G, T (nucleotide_symbols.h)
REFERENCE_SEED = 20200101
GENERATIONS = 5           # writeFullSequenceNdjson: SequenceTreeGenerator defaults
SPAN_DAYS = 1095          # dates 2020-01-01 .. 2022-12-30 spread evenly over the rows (sorted column)
FROM_DAY, TO_DAY = 366, 546  # 2021-01-01 .. 2021-06-30
MIN_PROPORTION = 0.05
METRIC = "mutations_query_seq_positions_per_s"
UNIT = "seq*positions/s"


def log(*args):
    print(*args, file=sys.stderr, flush=True)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as handle:
            return float(json.load(handle)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_bytes(args, algorithmic_bytes: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same command (profiles/<round>_k1_traffic.json, written by
    profiles/summarize_round.py); null when no capture of this workload is committed."""
    if args.traffic_bytes is not None:
        return args.traffic_bytes
    best = None
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
        if name.endswith("_k1_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as handle:
                    record = json.load(handle)
                if int(record["algorithmic_bytes_per_launch"]) == algorithmic_bytes:
                    best = float(record["dram_bytes_per_launch"])
            except Exception:
                continue
    return best


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.process = None
        self.lines = []

    def start(self):
        try:
            self.process = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._drain, daemon=True)
            self.thread.start()
        except Exception:
            self.process = None

    def _drain(self):
        for line in self.process.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.process is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.process.terminate()
        try:
            self.process.wait(timeout=2)
        except Exception:
            self.process.kill()
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
            except ValueError:
                continue
            for name, value in zip(names, parts[5:9]):
                if value.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(sm_max) if sm_max else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# the CPU side: oracle restatement of the reference path (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------

def lineage_row_ids(synthetic, ancestor: int, total_rows: int):
    """global row ids of the rows whose sequence descends from `ancestor` (row i holds evolved[i % n])"""
    import numpy as np
    n_sequences = synthetic.num_sequences
    in_lineage = np.zeros(n_sequences, dtype=bool)
    in_lineage[ancestor] = True
    for e in range(ancestor + 1, n_sequences):
        in_lineage[e] = in_lineage[synthetic.parent(e)]
    members = np.flatnonzero(in_lineage).astype(np.int64)
    ids = (np.arange(0, total_rows, n_sequences, dtype=np.int64)[:, None] + members[None, :]).ravel()
    return ids[ids < total_rows].astype(np.uint32)


def oracle_rows_that_fit(total_rows: int) -> int:
    """The oracle table of 10 M rows takes ~1.7 GB plus ~1.7 GB while it is imported; a host without the
    memory for the whole N-GPU table gets as many rows as fit (whole multiples of 10 M), and says so."""
    try:
        import psutil
        available = psutil.virtual_memory().available
    except Exception:
        return total_rows
    per_row = 3.6e9 / 1e7
    if total_rows * per_row * 1.25 <= available:
        return total_rows
    return max(10_000_000, int(available / 1.25 / per_row) // 10_000_000 * 10_000_000)


def build_oracle_table(total_rows: int, threads: int):
    """The oracle's table of THE SAME synthetic data the GPU arm queries, at full size: the product-side generator
    (host/synthetic.cpp; tests/test_host_cpu.py checks it against the oracle's own string generator) writes the
    column in the S1 upload format and the oracle imports it into its std::map of containers; the same lineage
    bitmap and the same per-chunk date ranges. Only data preparation happens here -- the timed query below
    runs on oracle code alone."""
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    started = time.perf_counter()
    synthetic = host_api.Synthetic(GENOME_LENGTH, REFERENCE_SEED, GENERATIONS)
    sizes = host_api.dense_chunk_sizes(total_rows)
    column = synthetic.build_column(total_rows, 0, len(sizes), threads)
    meta = {"containers": int(column.contents.n_containers), "payload_bytes": int(column.contents.payload_bytes)}
    table = O.Table()
    table.set_layout(*sizes)
    table.import_column("main", O.NUCLEOTIDE, synthetic.reference, column)
    synthetic.release_column()
    ancestor = next(e for e in range(synthetic.num_sequences) if synthetic.generation(e) == 2)
    table.register_bitmap("lineage", lineage_row_ids(synthetic, ancestor, total_rows))
    expression = (f"(and {host_api.date_ranges_expression(total_rows, SPAN_DAYS, FROM_DAY, TO_DAY, 0, len(sizes))} "
                  f"(bitmap lineage))")
    log(f"[oracle] table: {total_rows} rows, {meta['containers']} containers, {meta['payload_bytes'] / 1e9:.2f} GB payload, "
        f"built in {time.perf_counter() - started:.1f}s")
    return table, expression, meta


def time_oracle(table, expression, threads: int, min_seconds: float, max_queries: int):
    """threads concurrent independent native queries (the reference's one-worker-per-request model,
    no intra-query parallelism): computeFilter -> calculateMutationsPerPosition -> thresholding, with
    parsing/compiling inside the timer (performance/nof_sequence_filter.cpp:43-52 starts its clock in front of
    the planner). Python only starts the run and reads the totals."""
    queries, elapsed, cardinality = table.mutations_query_bench(
        "main", expression, MIN_PROPORTION, threads, min_seconds, max_queries)
    return cardinality * GENOME_LENGTH * queries / elapsed, elapsed, queries, cardinality


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_gpus = max(1, args.gpus)
    total_rows = args.rows_per_gpu * n_gpus
    rows = oracle_rows_that_fit(total_rows)
    table, expression, meta = build_oracle_table(rows, cores)
    n_output_rows, cardinality = table.mutations_query("main", expression, MIN_PROPORTION)  # warm-up
    single_value, single_elapsed, single_queries, _ = time_oracle(table, expression, 1, 2.0, 10 ** 9)
    per_step = []
    for step in range(args.warmup + args.steps):
        value, elapsed, queries, cardinality = time_oracle(table, expression, cores, args.reference_step_seconds, 10 ** 9)
        if step >= args.warmup:
            per_step.append((value, elapsed, queries))
    value = sum(v for v, _, _ in per_step) / len(per_step)
    ms_per_step = 1000.0 * sum(e for _, e, _ in per_step) / len(per_step)
    sample = (f"the whole table ({rows} rows x {GENOME_LENGTH} nt, same data, filter and minProportion as the GPU arm); "
              if rows == total_rows else
              f"{rows} of the {total_rows} rows (host memory), same generator, filter fractions and minProportion; ")
    sample += (f"a step = {cores} concurrent single-threaded queries running for >= {args.reference_step_seconds}s "
               f"({sum(q for _, _, q in per_step)} queries in the {len(per_step)} timed steps)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, rows, cardinality, n_output_rows, meta["containers"], meta["payload_bytes"]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "per_core_value": value / cores,
                         "single_thread": {"value": single_value, "cores": 1,
                                           "ms_per_query": 1000.0 * single_elapsed / max(1, single_queries)}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "run": {"parallelism": f"{cores} concurrent single-threaded queries on {cores} host cores (one worker per request, "
                               "api/api.cpp:39-51)", "timer": "parse + compile + filter + counts + thresholding inside"},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, total_rows, cardinality, output_rows, containers, payload_bytes):
    """What is computed, not how: identical in the GPU arm and the reference arm of one N."""
    return {
        "workload": "performance/mutation_benchmark-style synthetic full-length table, Mutations with date-range + lineage filter "
                    "(BASELINE.json configs[1])",
        "rows_per_gpu": args.rows_per_gpu, "total_rows": total_rows, "genome_length": GENOME_LENGTH,
        "min_proportion": MIN_PROPORTION,
        "filter": "date.between(2021-01-01, 2021-06-30) && lineage(generation-2 node, includeSublineages)",
        "filter_cardinality": cardinality, "output_rows": output_rows,
        "containers": containers, "payload_gb": round(payload_bytes / 1e9, 3),
        "l2": "the container payload touched per step exceeds the 126 MB L2, no explicit flush",
    }


# ---------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------

def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lapis_silo_b200 import abi, host_api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL may print its version banner to stdout while it builds the communicator; stdout must hold
        # the one JSON line only, so file descriptor 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.ones(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    n_gpus = world

    total_rows = args.rows_per_gpu * n_gpus
    sizes = host_api.dense_chunk_sizes(total_rows)
    # Partition scheduler: chunk c belongs to rank c % n_gpus. The date column is sorted, so the date
    # filter selects one contiguous row range; contiguous chunk ranges would hand it to one or two ranks.
    # An interleaved shard is a table of its own with local chunk ids (counts are plain addends).
    first, n_chunks, stride = host_api.interleaved_shard(len(sizes), n_gpus, rank)

    started = time.perf_counter()
    synthetic = host_api.Synthetic(GENOME_LENGTH, REFERENCE_SEED, GENERATIONS)
    ancestor = next(e for e in range(synthetic.num_sequences) if synthetic.generation(e) == 2)
    threads = max(1, (os.cpu_count() or 8) // max(1, min(n_gpus, 8)))
    column = synthetic.build_column(total_rows, first, n_chunks, threads, stride)
    payload_bytes = int(column.contents.payload_bytes)
    n_containers = int(column.contents.n_containers)
    built = time.perf_counter()
    ctx = abi.Context(local_rank)
    local_first = first if stride == 1 else 0
    table = host_api.HostTable(ctx, host_api.shard_chunk_sizes(total_rows, first, n_chunks, stride), first_chunk=local_first)
    table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, column)
    synthetic.release_column()
    table.register_bitmap("lineage", synthetic.lineage_bitmap(ancestor, total_rows, first, n_chunks, stride))
    expression = (f"(and {host_api.date_ranges_expression(total_rows, SPAN_DAYS, FROM_DAY, TO_DAY, first, n_chunks, stride)} "
                  f"(bitmap lineage))")
    torch.cuda.synchronize()
    uploaded = time.perf_counter()
    if rank == 0:
        log(f"[bench] rank 0 shard: chunks {first} + k*{stride}, k < {n_chunks}: {n_containers} containers, "
            f"{payload_bytes / 1e9:.2f} GB payload; generated in {built - started:.1f}s, uploaded in {uploaded - built:.1f}s; "
            f"{synthetic.num_sequences} evolved sequences")

    # a dedicated (non-default) stream: the library's kernels, NCCL and the timing events all use it
    valid_values = VALID_MUTATION_SYMBOLS * GENOME_LENGTH  # symbol ids 0..4 (-, A, C, G, T) are contiguous
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    counts = torch.zeros(N_SYMBOLS * GENOME_LENGTH, dtype=torch.int32, device="cuda")
    pinned = torch.zeros(N_SYMBOLS * GENOME_LENGTH, dtype=torch.int32).pin_memory()
    prepared = table.prepare(expression)

    # N > 1, default (--reduce peer): the library's own partition scheduler (silo_gpu_shard_group_*): the finalize kernel
    # of every rank stores its rows of the valid mutation symbols into rank 0's gather area over NVLink, rank 0's
    # own finalize kernel adds them and its counts -- no collective kernel, nothing but this repository's kernels on the stream. The
    # handles of the gather areas travel through torch.distributed once, at set-up.
    # --reduce nccl: the per-rank counts are all-reduced with NCCL instead; the all-reduce of query i runs on its own
    # stream beside the kernels of query i + 1 (two count buffers). Every reduction is inside the timed region.
    use_peer = (n_gpus > 1 or args.force_shard_group) and args.reduce == "peer"
    if use_peer:
        handles = [None] * n_gpus
        if n_gpus > 1:
            dist.all_gather_object(handles, table.shard_group_create("main", rank, n_gpus))
        else:  # (measurement only: what the sharded kernels cost without any peer)
            handles = [table.shard_group_create("main", 0, 1)]
        table.shard_group_connect(handles)
        if n_gpus > 1:
            dist.barrier()
    comm_stream = torch.cuda.Stream() if n_gpus > 1 and not use_peer else None
    count_buffers = [counts, torch.zeros_like(counts)] if n_gpus > 1 and not use_peer else [counts]
    kernels_done = [torch.cuda.Event() for _ in count_buffers]
    reduced = [torch.cuda.Event() for _ in count_buffers]
    issued = [0]

    def device_step():
        buffer = issued[0] % len(count_buffers)
        issued[0] += 1
        if use_peer:
            if rank == 0:  # filter + counts kernels; the finalize kernel waits for the other ranks' rows on the device and sums
                prepared.run_sharded_collect_async(stream.cuda_stream, counts.data_ptr())
            else:  # filter + counts kernels; the finalize kernel stores this rank's rows into rank 0's memory over NVLink
                prepared.run_sharded_async(stream.cuda_stream)
            return
        if n_gpus > 1:
            stream.wait_event(reduced[buffer])  # the all-reduce that used this buffer two queries ago
        prepared.run_counts_async(0, count_buffers[buffer].data_ptr(), stream.cuda_stream)  # filter + counts kernels
        if n_gpus > 1:
            kernels_done[buffer].record(stream)
            comm_stream.wait_event(kernels_done[buffer])
            with torch.cuda.stream(comm_stream):
                # the rows of the 5 valid mutation symbols (ids 0..4, contiguous): all that the action's output
                # pass reads (addMutationsToOutput, mutations_node.cpp:307-363); u32 viewed as i32, same bits
                dist.all_reduce(count_buffers[buffer][:valid_values])
                reduced[buffer].record(comm_stream)

    def join_reductions():
        for event in reduced if n_gpus > 1 and not use_peer else []:
            stream.wait_event(event)


    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(value: float) -> float:
        if n_gpus == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(value: int) -> int:
        if n_gpus == 1:
            return value
        t = torch.tensor([value], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return int(t.item())

    # ---- value: device-resident inputs ----
    # N = 1: the K steps of the timed region are captured into ONE CUDA graph (the launch gaps between the
    # dependent kernels of a query and the library's per-kernel timing events are ~12 % of an eager step).
    # Event records inside a capture are not replayed, so the roofline numbers come from a second, EAGER
    # pass over the same K steps right after the timed region: CUDA events on the launching stream around
    # every launch of the dominant kernel. N > 1: the same, the all-reduces (on their own stream) captured too.
    use_graph = not args.eager
    for _ in range(args.warmup):
        device_step()
    barrier()
    launches_before = table.stats().kernel_launches
    device_step()
    barrier()
    launches_per_step = int(table.stats().kernel_launches - launches_before)  # (this call also resets the per-kernel timing window)
    launches_before = table.stats().kernel_launches
    graph = None
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        if n_gpus == 1 or use_peer:
            with torch.cuda.graph(graph, stream=stream):
                for _ in range(args.steps):
                    device_step()
        else:
            # N > 1: the same two-stream pipeline, with events that live inside the capture (a captured
            # stream may only wait for work of the same capture) and the all-reduce stream joined at the end
            with torch.cuda.graph(graph, stream=stream):
                reduced_in_graph = []
                for index in range(args.steps):
                    buffer = index % len(count_buffers)
                    if index >= len(count_buffers):
                        stream.wait_event(reduced_in_graph[index - len(count_buffers)])
                    prepared.run_counts_async(0, count_buffers[buffer].data_ptr(), stream.cuda_stream)
                    done = torch.cuda.Event()
                    done.record(stream)
                    comm_stream.wait_event(done)
                    with torch.cuda.stream(comm_stream):
                        dist.all_reduce(count_buffers[buffer][:valid_values])
                        event = torch.cuda.Event()
                        event.record(comm_stream)
                    reduced_in_graph.append(event)
                for event in reduced_in_graph[-len(count_buffers):]:
                    stream.wait_event(event)
            issued[0] += args.steps
        graph.replay()  # instantiation / upload outside the timed region
        barrier()

    def timed_region():
        if graph is not None:
            graph.replay()
        else:
            for _ in range(args.steps):
                device_step()

    gpu_launches = launches_per_step * args.steps if graph is not None else None  # every replay launches what the capture recorded
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    begin.record(stream)
    timed_region()
    join_reductions()
    end.record(stream)
    barrier()
    own_device_ms = begin.elapsed_time(end)
    device_ms = max_over_ranks(own_device_ms)
    if n_gpus > 1:
        log(f"[bench] rank {rank}: {own_device_ms / args.steps * 1000:.1f} us per step on this rank's stream")
    if graph is not None:
        table.stats()  # reset the window: the eager pass below is what the per-kernel events describe
        for _ in range(args.steps):
            device_step()
        barrier()
    stats = table.stats()  # per-kernel CUDA events of exactly these launches
    if gpu_launches is None:
        gpu_launches = int(stats.kernel_launches - launches_before)
    # the timed region lasts a few milliseconds, nvidia-smi samples every 100 ms: keep the same step
    # running for another 0.4 s so that the clocks line describes the GPU under this load
    # (a step count, not a deadline: every rank must issue the same number of all-reduces)
    extra_steps = min(20000, max(args.steps, int(0.4 / max(device_ms / args.steps / 1000.0, 1e-6))))
    if graph is not None:
        for index in range(max(1, extra_steps // args.steps)):
            graph.replay()
            if index % 8 == 7:
                torch.cuda.synchronize()
    else:
        for index in range(extra_steps):
            device_step()
            if index % 64 == 63:
                torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled_over"] = "the timed region and 0.4 s of the same step right after it (nvidia-smi -lms 100)"
    if use_peer:  # (a sharded run forwards the filter's scalars to rank 0 and resets them: evaluate the filter once more)
        shard_filter = table.filter(expression)
        cardinality = sum_over_ranks(shard_filter.cardinality)
        shard_filter.close()
    else:
        cardinality = sum_over_ranks(prepared.cardinality())
    total_containers = sum_over_ranks(n_containers)
    total_payload_bytes = sum_over_ranks(payload_bytes)
    last_buffer = count_buffers[(issued[0] - 1) % len(count_buffers)]
    if n_gpus > 1 and not use_peer:  # outside the timed region: the other symbols' rows too, for the property check below
        dist.all_reduce(last_buffer[valid_values:])
        torch.cuda.synchronize()
    device_counts = last_buffer.cpu().numpy().view(np.uint32).reshape(N_SYMBOLS, GENOME_LENGTH).copy()
    if use_peer:  # rank 0 holds the summed rows of the valid symbols; the synthetic table has no other symbol
        device_counts[VALID_MUTATION_SYMBOLS:] = 0
    value = cardinality * GENOME_LENGTH * args.steps / (device_ms / 1000.0)

    # ---- e2e: host buffers in, host rows out, every step ----
    def e2e_step():
        # (the result comes back as columns, like the record batch the reference hands to its Arrow sink)
        if n_gpus == 1:
            return table.mutations_columns(["main"], expression, MIN_PROPORTION)  # MutationsNode through the C ABI
        # The sharded query (MutationsNode::enqueueShardCounts / collectRows): every rank parses + compiles the
        # query against its shard and enqueues program H2D + filter + counts, the counts of the valid symbols
        # are all-reduced on the same stream, rank 0 runs the output pass over the sums on the device and
        # gets the emitted tuples back; one host synchronisation per rank and step.
        if use_peer:
            # the library's scheduler: every rank enqueues (no host synchronisation on ranks > 0: a rank whose gather slot
            # has not been handed back waits on the device), rank 0 collects the rows of the whole table
            if rank == 0:
                return table.sharded_query("main", expression, MIN_PROPORTION)[0]  # enqueue + collect: one graph launch, one sync
            table.sharded_enqueue("main", expression, stream.cuda_stream)
            return None
        table.mutations_enqueue("main", expression, counts.data_ptr(), stream.cuda_stream)
        dist.all_reduce(counts[:valid_values])  # rows of the 5 valid symbols: all that the output pass reads
        if rank == 0:
            return table.mutations_collect("main", MIN_PROPORTION, counts.data_ptr(), stream.cuda_stream)[0]
        stream.synchronize()
        return None

    rows = None
    for _ in range(args.warmup):
        rows = e2e_step()
    barrier()
    wall = time.perf_counter()
    for _ in range(args.steps):
        rows = e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - wall) * 1000.0)
    e2e_value = cardinality * GENOME_LENGTH * args.steps / (e2e_ms / 1000.0)
    counts_bytes = N_SYMBOLS * GENOME_LENGTH * 4

    # size-independent parity properties at full size (tests/ hold the bit-exact oracle comparisons)
    column_sums = device_counts.sum(axis=0, dtype=np.uint64)
    if rank == 0 or not use_peer:
        assert (column_sums == cardinality).all(), "per-position symbol counts (summed over the ranks) must add up to the global |filter|"
    if rank == 0:
        rows = host_api.rows_from_columns(rows)
        # (device_counts: the all-reduced counts of the device-resident loop; thresholded on the host here)
        direct = table.mutation_rows_from_counts("main", device_counts, MIN_PROPORTION)
        assert direct == rows, "device-resident and host-buffer paths must emit identical rows"

    used_graph = graph is not None
    if n_gpus > 1:
        # release the captured graph (it holds NCCL kernels) before the communicator goes away, on every rank
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
    if rank != 0:
        dist.destroy_process_group()
        return
    peak, peak_source = measured_peak_gbs()
    kernel_ms = float(stats.last_counts_kernel_ms)
    achieved = stats.counts_kernel_bytes / (kernel_ms / 1000.0) / 1e9 if kernel_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": device_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, total_rows, cardinality, len(rows), total_containers, total_payload_bytes),
        "run": {
            "chunks_per_gpu": n_chunks, "containers_per_gpu": n_containers, "payload_gb_per_gpu": round(payload_bytes / 1e9, 3),
            "launch": f"one CUDA graph holding the {args.steps} steps of the timed region" + (", all-reduces included" if n_gpus > 1 else "")
            if used_graph else "eager launches",
            "parallelism": (f"interleaved chunk shards (chunk c on rank c % {n_gpus}), " + (
                "the library's shard group: every rank's finalize kernel stores its rows into rank 0's gather area over NVLink, "
                "rank 0's finalize kernel adds them to its own (no collective kernel)" if use_peer else "NCCL allreduce of the u32 counts"))
            if n_gpus > 1 else "single GPU",
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": prepared.staged_bytes,
                # N = 1: the output pass runs on the device; the finalize kernel stores the emitted (position,
                # symbol, count, total) tuples and a 16-byte header (tuple count, filter cardinality, error
                # flag) straight into page-locked host memory (silo_gpu_query_mutation_hits);
                # at every N (N > 1: on rank 0, after the all-reduce of the counts on the device)
                "d2h_bytes_per_step": (len(rows) + 1) * 16},
        "gpu_launches": gpu_launches,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic_bytes(args, int(stats.counts_kernel_bytes)), "kernel": "containerAndCountKernel",
            "algorithmic_bytes_per_launch": int(stats.counts_kernel_bytes), "kernel_ms": kernel_ms,
            "timed_launches": int(stats.timed_calls), "peak_source": peak_source,
            "timed_with": "CUDA events on the launching stream around every launch, " + (
                "eager pass over the same steps right after the graph-timed region" if used_graph else "inside the timed region"),
            "whole_query_algorithmic_bytes": int(stats.algorithmic_bytes), "whole_query_ms": float(stats.last_total_ms),
        },
    }
    if n_gpus == 1 and not args.skip_cpu_baseline:
        # The oracle on THE SAME table at full size: (1) the bit-exact check of this run's results -- counts of all
        # 16 symbols at all 29,903 positions, the filtered row count and the thresholded output rows with their
        # double proportions --, (2) the CPU baseline, one thread per query like the reference.
        cores = os.cpu_count() or 1
        oracle_table, oracle_expression, _ = build_oracle_table(total_rows, cores)
        oracle_filter = oracle_table.filter(oracle_expression)
        assert oracle_filter.cardinality == cardinality, (oracle_filter.cardinality, cardinality)
        oracle_counts = oracle_table.mutation_counts("main", oracle_filter)
        assert np.array_equal(oracle_counts, device_counts), "device counts differ from the oracle's at full size"
        oracle_rows = oracle_table.mutation_rows("main", oracle_counts, MIN_PROPORTION)
        assert oracle_rows == rows, "output rows differ from the oracle's at full size"
        line["parity"] = {"oracle": "full size, same table", "rows": total_rows, "filter_cardinality_equal": True,
                          "counts_equal": True, "output_rows_equal": True, "output_rows": len(rows)}
        cpu_value, cpu_elapsed, cpu_queries, cpu_cardinality = time_oracle(
            oracle_table, oracle_expression, 1, args.cpu_seconds, 10 ** 9)
        all_value, all_elapsed, all_queries, _ = time_oracle(oracle_table, oracle_expression, cores, args.cpu_seconds / 2, 10 ** 9)
        line["cpu_baseline"] = {
            "value": cpu_value, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"the whole table ({total_rows} rows x {GENOME_LENGTH} nt, the GPU arm's data, filter and minProportion); "
                      f"{cpu_queries} single-threaded queries in {cpu_elapsed:.1f}s ({1000.0 * cpu_elapsed / cpu_queries:.0f} ms/query, "
                      f"|filter| = {cpu_cardinality}); host has {cores} cores",
            "all_cores": {"value": all_value, "cores": cores,
                          "sample": f"{all_queries} queries by {cores} concurrent single-threaded workers in {all_elapsed:.1f}s"},
        }
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if n_gpus > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# --workload nof: BASELINE.json configs[2], performance/nof_sequence_filter.cpp:69-80,150-178
# ---------------------------------------------------------------------------------------------

NOF_DISTANCES = (0, 5, 50, 200)  # nof_sequence_filter.cpp:164
NOF_METRIC = "nof_filter_rows_per_s"
NOF_UNIT = "rows/s"
NOF_CPU_SAMPLE_ROWS = 2 * 65536  # the oracle's Threshold DP is O(n k) whole-bitmap passes: 541 s for distance 50 at 10 M rows


def nof_config(args, total_rows, cardinalities):
    return {
        "workload": "performance/nof_sequence_filter: nucleotideMutationProfile(distance, querySequence = last evolved sequence) -> count() "
                    "on the mutation_benchmark-style full-length table (BASELINE.json configs[2]); step i uses distance "
                    f"{list(NOF_DISTANCES)}[i % 4]",
        "rows_per_gpu": args.rows_per_gpu, "total_rows": total_rows, "genome_length": GENOME_LENGTH,
        "children_per_query": GENOME_LENGTH, "filter_cardinalities": cardinalities,
        "l2": "every query streams the column's whole container payload (1.42 GB per GPU), no explicit flush",
    }


def nof_oracle_sample(threads: int):
    from lapis_silo_b200 import host_api
    from oracle import oracle as O
    synthetic = host_api.Synthetic(GENOME_LENGTH, REFERENCE_SEED, GENERATIONS)
    sizes = host_api.dense_chunk_sizes(NOF_CPU_SAMPLE_ROWS)
    table = O.Table()
    table.set_layout(*sizes)
    table.import_column("main", O.NUCLEOTIDE, synthetic.reference, synthetic.build_column(NOF_CPU_SAMPLE_ROWS, 0, len(sizes), threads))
    synthetic.release_column()
    return table, synthetic.sequence(synthetic.num_sequences - 1)


def nof_time_oracle(table, query, distances, threads: int):
    """threads concurrent workers, each runs the queries of `distances` once (parse + rewrite + compile + evaluate + count
    inside the timer, performance/nof_sequence_filter.cpp:43-52); returns (rows/s, seconds, queries)."""
    from concurrent.futures import ThreadPoolExecutor

    def worker(_):
        for distance in distances:
            flt = table.filter(f"(profile main {distance} seq {query})")
            flt.cardinality
            flt.close()
    started = time.perf_counter()
    with ThreadPoolExecutor(threads) as pool:  # (ctypes releases the GIL inside the oracle)
        list(pool.map(worker, range(threads)))
    elapsed = time.perf_counter() - started
    queries = threads * len(distances)
    return NOF_CPU_SAMPLE_ROWS * queries / elapsed, elapsed, queries


def run_nof_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    table, query = nof_oracle_sample(cores)
    per_step = []
    for step in range(args.warmup + args.steps):
        distance = NOF_DISTANCES[step % len(NOF_DISTANCES)]
        value, elapsed, queries = nof_time_oracle(table, query, [distance], cores)
        if step >= args.warmup:
            per_step.append((elapsed, queries))
    seconds = sum(e for e, _ in per_step)
    queries = sum(q for _, q in per_step)
    value = NOF_CPU_SAMPLE_ROWS * queries / seconds
    total_rows = args.rows_per_gpu * max(1, args.gpus)
    sample = (f"{NOF_CPU_SAMPLE_ROWS} rows of the same table (the oracle's Threshold DP needs 541 s for distance 50 at 10 M rows); a step = "
              f"{cores} concurrent single-threaded queries of the step's distance")
    print(json.dumps({
        "impl": "reference", "metric": NOF_METRIC, "value": value, "unit": NOF_UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * seconds / len(per_step), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16",
        "data": "synthetic", "config": nof_config(args, total_rows, None),
        "cpu_baseline": {"value": value, "unit": NOF_UNIT, "cores": cores, "kind": "port", "sample": sample, "per_core_value": value / cores},
        "e2e": {"value": value, "unit": NOF_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }), flush=True)


def run_nof(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lapis_silo_b200 import abi, host_api
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("gloo")  # filters need no exchange (row-local): only the timings and cardinalities meet
    n_gpus = world
    total_rows = args.rows_per_gpu * n_gpus
    sizes = host_api.dense_chunk_sizes(total_rows)
    first, n_chunks, stride = host_api.interleaved_shard(len(sizes), n_gpus, rank)
    synthetic = host_api.Synthetic(GENOME_LENGTH, REFERENCE_SEED, GENERATIONS)
    threads = max(1, (os.cpu_count() or 8) // max(1, min(n_gpus, 8)))
    ctx = abi.Context(local_rank)
    table = host_api.HostTable(ctx, host_api.shard_chunk_sizes(total_rows, first, n_chunks, stride), first_chunk=first if stride == 1 else 0)
    table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, first, n_chunks, threads, stride))
    synthetic.release_column()
    query = synthetic.sequence(synthetic.num_sequences - 1)
    texts = [f"(profile main {distance} seq {query})" for distance in NOF_DISTANCES]
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def reduce(value, op):
        if n_gpus == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def barrier():
        torch.cuda.synchronize()
        if n_gpus > 1:
            dist.barrier()

    # ---- value: the programs are device resident, a step enqueues the sweep and the interpreter ----
    prepared = [table.prepare(text) for text in texts]
    for step in range(args.warmup):
        prepared[step % len(prepared)].run_async(stream.cuda_stream)
    barrier()
    table.sweep_stats()
    launches_before = table.stats().kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    begin.record(stream)
    for step in range(args.steps):
        prepared[step % len(prepared)].run_async(stream.cuda_stream)
    end.record(stream)
    barrier()
    device_ms = reduce(begin.elapsed_time(end), dist.ReduceOp.MAX if n_gpus > 1 else None)
    gpu_launches = int(table.stats().kernel_launches - launches_before)
    sweep_ms, sweep_bytes, sweep_calls = table.sweep_stats()
    # (the clocks line: keep the same step running for 0.4 s)
    for step in range(max(args.steps, int(0.4 / max(device_ms / args.steps / 1000.0, 1e-6)))):
        prepared[step % len(prepared)].run_async(stream.cuda_stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    cardinalities = [int(reduce(p.cardinality(), dist.ReduceOp.SUM if n_gpus > 1 else None)) for p in prepared]
    value = total_rows * args.steps / (device_ms / 1000.0)

    # ---- e2e: expression text in, count out, through the host layer, every step ----
    def e2e_step(step):
        flt = table.filter(texts[step % len(texts)])
        count = flt.cardinality
        flt.close()
        return count
    for step in range(args.warmup):
        e2e_step(step)
    barrier()
    wall = time.perf_counter()
    counts = [e2e_step(step) for step in range(args.steps)]
    barrier()
    e2e_ms = reduce((time.perf_counter() - wall) * 1000.0, dist.ReduceOp.MAX if n_gpus > 1 else None)
    lowered = table.lower_timed(texts[1])
    if rank != 0:
        return
    peak, peak_source = measured_peak_gbs()
    achieved = sweep_bytes / (sweep_ms / 1000.0) / 1e9 if sweep_ms > 0 else 0.0
    line = {
        "metric": NOF_METRIC, "value": value, "unit": NOF_UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": device_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": nof_config(args, total_rows, cardinalities),
        "run": {"launch": "eager launches (sweep + interpreter per query)", "parallelism": f"interleaved chunk shards, no exchange (filters are row-local)" if n_gpus > 1 else "single GPU"},
        "clocks": clocks,
        "e2e": {"value": total_rows * args.steps / (e2e_ms / 1000.0), "unit": NOF_UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(lowered["blob_bytes"]) + 16 * int(lowered["n_instrs"]), "d2h_bytes_per_step": 16,
                "host_lowering_us": lowered["parse_us"] + lowered["rewrite_us"] + lowered["compile_us"] + lowered["lower_us"]},
        "gpu_launches": gpu_launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": args.traffic_bytes,
                     "kernel": "thresholdSweepKernel", "algorithmic_bytes_per_launch": sweep_bytes, "kernel_ms": sweep_ms, "timed_launches": sweep_calls,
                     "peak_source": peak_source, "timed_with": "CUDA events on the launching stream around every launch of the timed region"},
    }
    if n_gpus == 1 and not args.skip_cpu_baseline:
        # full-size check of this run's row counts against the brute-force Hamming distances (tests/ pin it to the oracle's DP)
        n = synthetic.num_sequences
        sequences = np.array([np.frombuffer(synthetic.sequence(e).encode(), dtype=np.uint8) for e in range(n)])
        distances = (sequences != sequences[n - 1]).sum(axis=1)
        row_counts = np.bincount(np.arange(total_rows) % n, minlength=n)
        want = [int(row_counts[distances <= d].sum()) for d in NOF_DISTANCES]
        assert cardinalities == want and counts[:len(NOF_DISTANCES)] == want, (cardinalities, want)
        line["parity"] = {"checked_against": "brute-force Hamming distances at full size (pinned to the oracle's Threshold DP in tests/)", "cardinalities_equal": True}
        oracle_table, oracle_query = nof_oracle_sample(os.cpu_count() or 1)
        cpu_value, cpu_elapsed, cpu_queries = nof_time_oracle(oracle_table, oracle_query, NOF_DISTANCES, 1)
        line["cpu_baseline"] = {"value": cpu_value, "unit": NOF_UNIT, "cores": 1, "kind": "port",
                                "sample": f"{NOF_CPU_SAMPLE_ROWS} rows of the same table, the four distances once each: {cpu_elapsed:.1f}s single-threaded "
                                          "(the oracle's Threshold DP is O(n k): 9 / 59 / 541 s for distance 0 / 5 / 50 at 10 M rows)"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# --workload aa: BASELINE.json configs[3], AminoAcidMutations over the 12 SARS-CoV-2 genes
# ---------------------------------------------------------------------------------------------

GENE_LENGTHS = {"E": 76, "M": 223, "N": 420, "ORF1a": 4401, "ORF1b": 2696, "ORF3a": 276, "ORF6": 62, "ORF7a": 122, "ORF7b": 44,
                "ORF8": 122, "ORF9b": 98, "S": 1274}  # testBaseData/exampleDataset/reference_genomes.json; sum = 9,814
AA_MUTATION_RATE = 0.003  # SURVEY.md 8(d) input 4: the tree model per gene over the valid amino-acid symbols, mu = 0.001 * 3
AA_METRIC = "aa_mutations_query_seq_positions_per_s"


def aa_genes():
    from lapis_silo_b200 import host_api
    return {name: host_api.Synthetic(genome_length=length, reference_seed=REFERENCE_SEED + index, generations=GENERATIONS, gene=True,
                                     tree_seed=42 + index, mutation_rate=AA_MUTATION_RATE)
            for index, (name, length) in enumerate(GENE_LENGTHS.items())}


def aa_lineage_rows(genes, total_rows):
    """the config-2 lineage filter needs a tree: the rows that descend from a generation-2 node of ORF1a's tree"""
    tree = genes["ORF1a"]
    ancestor = next(e for e in range(tree.num_sequences) if tree.generation(e) == 2)
    return ancestor, lineage_row_ids(tree, ancestor, total_rows)


def aa_config(args, total_rows, cardinality, output_rows):
    return {
        "workload": "AminoAcidMutations over the 12 SARS-CoV-2 genes (sum of lengths 9,814; tree model per gene over the valid amino-acid symbols, "
                    "mu = 0.003), config-2 filter (date range + lineage), minProportion 0.05 (BASELINE.json configs[3])",
        "rows_per_gpu": args.rows_per_gpu, "total_rows": total_rows, "genes": len(GENE_LENGTHS), "positions": sum(GENE_LENGTHS.values()),
        "min_proportion": MIN_PROPORTION, "filter_cardinality": cardinality, "output_rows": output_rows,
    }


def run_aa(args):
    import numpy as np
    import torch
    from lapis_silo_b200 import abi, host_api
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("--workload aa is a single-GPU line (N > 1: the Mutations workload is the scaling line)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(0)
    total_rows = args.rows_per_gpu
    sizes = host_api.dense_chunk_sizes(total_rows)
    threads = os.cpu_count() or 8
    genes = aa_genes()
    ctx = abi.Context(0)
    table = host_api.HostTable(ctx, sizes)
    for name, gene in genes.items():
        column = gene.build_column(total_rows, 0, len(sizes), threads)
        table.add_column(name, host_api.AMINO_ACID, gene.reference, column)
        gene.release_column()
    ancestor, lineage = aa_lineage_rows(genes, total_rows)
    table.register_bitmap("lineage", genes["ORF1a"].lineage_bitmap(ancestor, total_rows, 0, len(sizes)))
    expression = f"(and {host_api.date_ranges_expression(total_rows, SPAN_DAYS, FROM_DAY, TO_DAY, 0, len(sizes))} (bitmap lineage))"
    names = list(GENE_LENGTHS)
    positions = sum(GENE_LENGTHS.values())

    def step():
        return table.mutations_columns(names, expression, MIN_PROPORTION)  # MutationsNode: one device call, one sync
    for _ in range(args.warmup):
        rows = step()
    torch.cuda.synchronize()
    launches_before = table.stats().kernel_launches
    sampler = ClockSampler(0)
    sampler.start()
    wall = time.perf_counter()
    for _ in range(args.steps):
        rows = step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - wall) * 1000.0
    gpu_launches = int(table.stats().kernel_launches - launches_before)
    for _ in range(int(0.4 / max(e2e_ms / args.steps / 1000.0, 1e-6))):
        step()
    clocks = sampler.stop()
    flt = table.filter(expression)
    cardinality = flt.cardinality
    flt.close()
    # algorithmic bytes of one query: per gene, descriptors + payloads of the containers of the chunks that
    # hold a filtered row + their filter tiles, as the library accounts them for the container kernel
    prepared = table.prepare(expression)
    scratch = torch.zeros(28 * max(GENE_LENGTHS.values()), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()
    prepared.run_async(stream.cuda_stream)
    payload_bytes = 0
    for index in range(len(names)):
        table.mutation_counts_async(index, prepared, scratch.data_ptr(), stream.cuda_stream)
        payload_bytes += int(table.stats().counts_kernel_bytes)
    prepared.close()
    rows = host_api.rows_from_columns(rows)
    value = cardinality * positions * args.steps / (e2e_ms / 1000.0)
    peak, peak_source = measured_peak_gbs()
    line = {
        "metric": AA_METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": e2e_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": aa_config(args, total_rows, cardinality, len(rows)),
        "run": {"launch": "one replayed CUDA graph per query: the filter program once, then work list / coverage / container / finalize kernels per gene",
                "note": "value is measured through the host API like e2e: the query is launch-latency bound (49 small kernels), not HBM bound"},
        "clocks": clocks,
        "e2e": {"value": value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": 1472, "d2h_bytes_per_step": (len(rows) + len(names)) * 16},
        "gpu_launches": gpu_launches,
        "roofline": {"bound": "hbm", "achieved": payload_bytes * (cardinality > 0) / (e2e_ms / args.steps / 1000.0) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": payload_bytes / (e2e_ms / args.steps / 1000.0) / 1e9 / peak, "traffic": None, "kernel": "whole query (12 x containerAndCountKernel + small kernels)",
                     "algorithmic_bytes_per_launch": payload_bytes, "kernel_ms": e2e_ms / args.steps, "peak_source": peak_source,
                     "note": "bytes: the container kernels' algorithmic bytes summed over the 12 genes; time: the whole query through the host API"},
    }
    if not args.skip_cpu_baseline:
        from oracle import oracle as O
        oracle_table = O.Table()
        oracle_table.set_layout(*sizes)
        for name, gene in genes.items():
            oracle_table.import_column(name, O.AMINO_ACID, gene.reference, gene.build_column(total_rows, 0, len(sizes), threads))
            gene.release_column()
        oracle_table.register_bitmap("lineage", lineage)
        started = time.perf_counter()
        want = [row for name in names for row in oracle_table.mutations(name, expression, MIN_PROPORTION)]
        assert want == rows, "output rows differ from the oracle's at full size"
        line["parity"] = {"oracle": "full size, same table", "output_rows_equal": True, "output_rows": len(rows)}
        queries, seconds = 0, 0.0
        while seconds < args.cpu_seconds:
            started = time.perf_counter()
            flt = oracle_table.filter(expression)
            for name in names:
                oracle_table.mutation_rows(name, oracle_table.mutation_counts(name, flt), MIN_PROPORTION)
            flt.close()
            seconds += time.perf_counter() - started
            queries += 1
        line["cpu_baseline"] = {"value": cardinality * positions * queries / seconds, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"the whole table, {queries} single-threaded queries in {seconds:.1f}s (filter once, then the 12 genes)"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# --workload reads: BASELINE.json configs[2], second half: performance/many_short_read_filters.cpp
# ---------------------------------------------------------------------------------------------

READS_COUNT, READS_LENGTH, READS_QUERIES = 5_000_000, 200, 10_000  # sequence_generator.h:189-190, many_short_read_filters.cpp:25
READS_METRIC = "many_short_read_filters_queries_per_s"
READS_DAY = 19723  # 2024-01-01 as Date32


def reads_queries(genome_length: int, n_queries: int):
    """many_short_read_filters.cpp:42-87: a random 1-based position, alternately one symbol of (A, C, G, T, -) and the Or of
    all five, under `locationName = 'generated'` and the same samplingDate range twice. (numpy's generator instead of
    std::mt19937 + uniform_int_distribution: other positions, the same distribution; both arms run the same list.)"""
    rng = np_rng(42)
    symbols = "ACGT-"
    queries = []
    for index in range(n_queries):
        position = int(rng.integers(1, genome_length))
        date = f"(date-between samplingDate {READS_DAY} {READS_DAY + 6})"
        if index % 2 == 1:
            sequence = "(or " + " ".join(f"(sym-eq main {position} {symbol})" for symbol in symbols) + ")"
        else:
            sequence = f"(sym-eq main {position} {symbols[int(rng.integers(0, 5))]})"
        queries.append(f"(and (str-eq locationName generated) {date} {sequence} {date})")
    return queries


def np_rng(seed):
    import numpy as np
    return np.random.default_rng(seed)


def run_reads(args):
    import numpy as np
    import torch
    from lapis_silo_b200 import abi, host_api
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("--workload reads is a single-GPU line")
    reference_arm = args.impl == "reference"
    if not reference_arm and not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    threads = os.cpu_count() or 8
    synthetic = host_api.Synthetic(GENOME_LENGTH, REFERENCE_SEED, GENERATIONS)
    synthetic.draw_short_reads(READS_COUNT, READS_LENGTH)
    sizes = host_api.dense_chunk_sizes(READS_COUNT)
    queries = reads_queries(GENOME_LENGTH, READS_QUERIES)
    ids = np.zeros(READS_COUNT, dtype=np.uint32)
    days = np.full(READS_COUNT, READS_DAY, dtype=np.int32)
    config = {
        "workload": f"performance/many_short_read_filters: {READS_QUERIES} count() queries (a symbol test or a five-symbol Or at a random position, "
                    "AND locationName = 'generated' AND samplingDate between, twice) over "
                    f"{READS_COUNT} short reads x {READS_LENGTH} nt tiled over a {GENOME_LENGTH}-nt genome (BASELINE.json configs[2])",
        "rows": READS_COUNT, "read_length": READS_LENGTH, "queries": READS_QUERIES, "genome_length": GENOME_LENGTH,
    }

    def oracle_table():
        from oracle import oracle as O
        table = O.Table()
        table.set_layout(*sizes)
        table.import_column("main", O.NUCLEOTIDE, synthetic.reference, synthetic.build_short_read_column(0, len(sizes), threads))
        synthetic.release_column()
        table.add_string_column_ids("locationName", ["generated"], ids)
        table.add_date_column("samplingDate", days)
        return table

    def oracle_run(table, subset):
        started = time.perf_counter()
        counts = []
        for text in subset:
            flt = table.filter(text)
            counts.append(flt.cardinality)
            flt.close()
        return time.perf_counter() - started, counts

    if reference_arm:
        table = oracle_table()
        sample = queries[::50]  # 200 of the 10,000 queries
        oracle_run(table, sample[:5])
        seconds, _ = oracle_run(table, sample)
        value = len(sample) / seconds
        print(json.dumps({
            "impl": "reference", "metric": READS_METRIC, "value": value, "unit": "queries/s", "n_gpus": 1, "steps": 1, "warmup": 0,
            "ms_per_step": seconds * 1000.0, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": 1, "kind": "port",
                             "sample": f"every 50th of the {READS_QUERIES} queries ({len(sample)}) on the whole table, one thread: {seconds:.1f}s; "
                                       f"all {READS_QUERIES} would take ~{seconds * 50:.0f}s"},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
        }), flush=True)
        return

    torch.cuda.set_device(0)
    ctx = abi.Context(0)
    table = host_api.HostTable(ctx, sizes)
    table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_short_read_column(0, len(sizes), threads))
    synthetic.release_column()
    table.add_string_column_ids("locationName", ["generated"], ids)
    table.add_date_column("samplingDate", days)

    def run(subset):
        return [table.count(text) for text in subset]  # CountFilterNode through the host layer: one device call per query
    run(queries[:200])  # warm-up
    torch.cuda.synchronize()
    launches_before = table.stats().kernel_launches
    sampler = ClockSampler(0)
    sampler.start()
    started = time.perf_counter()
    counts = run(queries)
    seconds = time.perf_counter() - started
    clocks = sampler.stop()
    gpu_launches = int(table.stats().kernel_launches - launches_before)
    value = len(queries) / seconds
    peak, peak_source = measured_peak_gbs()
    # algorithmic bytes of one query ("NOf / SymbolInSet filter" row, plus the predicate's column): 4 B per row for
    # the string predicate, 8 B per row of (start, end) when the symbol set holds the local reference (about every other
    # query), the result tiles
    bytes_per_query = 4 * READS_COUNT + 8 * READS_COUNT // 2 + 8192 * len(sizes)
    line = {
        "metric": READS_METRIC, "value": value, "unit": "queries/s", "n_gpus": 1, "steps": 1, "warmup": 0, "ms_per_step": seconds * 1000.0,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
        "run": {"launch": "one filter program per query through the host API (parse, compile, lower, H2D, interpreter kernel, count back)",
                "seconds_for_all_queries": seconds, "us_per_query": seconds / len(queries) * 1e6},
        "clocks": clocks,
        "e2e": {"value": value, "unit": "queries/s", "ms_per_step": seconds * 1000.0, "h2d_bytes_per_step": 600 * len(queries), "d2h_bytes_per_step": 16 * len(queries)},
        "gpu_launches": gpu_launches,
        "roofline": {"bound": "hbm", "achieved": bytes_per_query * len(queries) / seconds / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bytes_per_query * len(queries) / seconds / 1e9 / peak, "traffic": None, "kernel": "evalProgramKernel (whole query, host time included)",
                     "algorithmic_bytes_per_launch": bytes_per_query, "kernel_ms": seconds / len(queries) * 1000.0, "peak_source": peak_source},
    }
    if not args.skip_cpu_baseline:
        oracle = oracle_table()
        sample_index = list(range(0, len(queries), 50))
        oracle_seconds, oracle_counts = oracle_run(oracle, [queries[i] for i in sample_index])
        assert oracle_counts == [counts[i] for i in sample_index], "query counts differ from the oracle's"
        line["parity"] = {"oracle": "full size, same table", "queries_compared": len(sample_index), "counts_equal": True}
        line["cpu_baseline"] = {"value": len(sample_index) / oracle_seconds, "unit": "queries/s", "cores": 1, "kind": "port",
                                "sample": f"every 50th query ({len(sample_index)}) on the whole table, one thread: {oracle_seconds:.1f}s "
                                          f"(all {len(queries)}: ~{oracle_seconds * 50:.0f}s)"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)


COOC_SEQUENCES = 2_000_000  # performance/sequence_generator.h:488-490
COOC_POSITIONS = (5, 10, 20, 30, 40, 50)  # 1-based, co_occurrence_benchmark.cpp:41
COOC_METRIC = "co_occurrence_rows_per_s"


def run_cooc(args):
    """BASELINE.json configs[4], second half: performance/co_occurrence_benchmark.cpp -- map({s_i := main.at(p_i)}).groupBy(count)
    over six positions of the 2 M x 100-nt random table (cycled to --rows-per-gpu rows per GPU: x 5 at the default, x 40 over
    8 GPUs). One process per GPU: every rank aggregates its interleaved chunk shard on the device
    (BitmapAggregationNode::executeShard -> silo_gpu_query_combinations), the (key, count) lists meet on rank 0
    (all_gather of a fixed 4,097 x 2 u64 tensor), rank 0 sums them per key and materialises the rows (::mergeShards)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from lapis_silo_b200 import abi, host_api
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    reference_arm = args.impl == "reference"
    if not reference_arm and not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    n_gpus = world
    total_rows = args.rows_per_gpu * n_gpus
    sizes = host_api.dense_chunk_sizes(total_rows)
    threads = max(1, (os.cpu_count() or 8) // max(1, min(n_gpus, 8)))
    dimensions = [("position", "main", p - 1) for p in COOC_POSITIONS]
    config = {
        "workload": f"performance/co_occurrence_benchmark: map(s_i := main.at(p_i)).groupBy(count) over the positions {list(COOC_POSITIONS)} of the "
                    f"{COOC_SEQUENCES} x 100-nt random table (Binomial(100, 0.1) substitutions per row), cycled to {total_rows} rows (BASELINE.json configs[4])",
        "rows_per_gpu": args.rows_per_gpu, "total_rows": total_rows, "genome_length": 100, "positions": list(COOC_POSITIONS),
    }
    if reference_arm and rank != 0:
        return
    synthetic = host_api.Synthetic(co_occurrence_sequences=COOC_SEQUENCES)

    def oracle_table(n_rows):
        from oracle import oracle as O
        table = O.Table()
        chunk_sizes = host_api.dense_chunk_sizes(n_rows)
        table.set_layout(*chunk_sizes)
        table.import_column("main", O.NUCLEOTIDE, synthetic.reference, synthetic.build_column(n_rows, 0, len(chunk_sizes), os.cpu_count() or 8))
        synthetic.release_column()
        return table

    if reference_arm:
        # the reference's own run: one thread per query (the recursive partition of bitmap_aggregation_node.cpp:91-116 is sequential)
        sample_rows = min(total_rows, COOC_SEQUENCES)
        table = oracle_table(sample_rows)
        table.bitmap_aggregation(dimensions, None)
        times = []
        for _ in range(max(1, args.steps)):
            started = time.perf_counter()
            rows = table.bitmap_aggregation(dimensions, None)
            times.append(time.perf_counter() - started)
        seconds = sum(times) / len(times)
        value = sample_rows / seconds
        print(json.dumps({
            "impl": "reference", "metric": COOC_METRIC, "value": value, "unit": "rows/s", "n_gpus": n_gpus, "steps": len(times), "warmup": 1,
            "ms_per_step": seconds * 1000.0, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "rows/s", "cores": 1, "kind": "port",
                             "sample": f"the benchmark's own {sample_rows} rows (one cycle of the table), {len(times)} queries, one thread: {seconds * 1000:.0f} ms per query, "
                                       f"{len(rows)} combinations"},
            "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
        }), flush=True)
        return

    torch.cuda.set_device(local_rank)
    if n_gpus > 1:
        dist.init_process_group("nccl")
    first, n_chunks, stride = host_api.interleaved_shard(len(sizes), n_gpus, rank)
    ctx = abi.Context(local_rank)
    table = host_api.HostTable(ctx, host_api.shard_chunk_sizes(total_rows, first, n_chunks, stride), first_chunk=first if stride == 1 else 0)
    table.add_column("main", host_api.NUCLEOTIDE, synthetic.reference, synthetic.build_column(total_rows, first, n_chunks, threads, stride))
    synthetic.release_column()
    max_entries = 4096  # 4^6 combinations of A/C/G/T: the table holds no other symbol
    mine = torch.zeros((max_entries + 1, 2), dtype=torch.int64, device="cuda")
    gathered = torch.zeros((n_gpus, max_entries + 1, 2), dtype=torch.int64, device="cuda")
    staging = torch.zeros((max_entries + 1, 2), dtype=torch.int64).pin_memory()

    def query():
        if n_gpus == 1:
            return table.bitmap_aggregation_columns(dimensions, None)  # BitmapAggregationNode::execute, the rows as arrays
        pairs, cardinality = table.bitmap_aggregation_shard(dimensions, None)  # the device call of this rank's shard; synchronous
        staging[0, 0], staging[0, 1] = len(pairs), cardinality
        staging[1:1 + len(pairs)] = torch.from_numpy(pairs.view(np.int64))
        mine.copy_(staging, non_blocking=True)
        dist.all_gather_into_tensor(gathered, mine)
        if rank != 0:
            return None
        host = gathered.cpu().numpy().view(np.uint64)
        return table.bitmap_aggregation_merge_columns(dimensions, [(host[r, 1:1 + int(host[r, 0, 0])], int(host[r, 0, 1])) for r in range(n_gpus)])

    def barrier():
        torch.cuda.synchronize()
        if n_gpus > 1:
            dist.barrier()

    def reduce_max(value):
        if n_gpus == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        query()
    # ---- value: the device call alone (every rank's shard, no exchange), CUDA events around the K calls ----
    barrier()
    launches_before = table.stats().kernel_launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    begin.record()
    for _ in range(args.steps):
        table.bitmap_aggregation_shard(dimensions, None)
    end.record()
    barrier()
    device_ms = reduce_max(begin.elapsed_time(end))
    gpu_launches = int(table.stats().kernel_launches - launches_before)
    # ---- e2e: the whole query -- device call, the lists to rank 0, merge, rows ----
    barrier()
    wall = time.perf_counter()
    for _ in range(args.steps):
        result = query()
    torch.cuda.synchronize()
    e2e_ms = reduce_max((time.perf_counter() - wall) * 1000.0)
    rows = host_api.combination_rows_from_columns(*result) if result is not None else None
    clocks = sampler.stop()
    if rank != 0:
        dist.destroy_process_group()
        return
    peak, peak_source = measured_peak_gbs()
    # algorithmic bytes of one query on one rank (co-occurrence row): one code byte per row and dimension
    # written and read back, plus the rows' (start, end) pairs that decide "reference symbol or missing"
    rows_here = sum(host_api.shard_chunk_sizes(total_rows, first, n_chunks, stride))
    bytes_per_query = rows_here * (2 * len(dimensions) + 8)
    line = {
        "metric": COOC_METRIC, "value": total_rows * args.steps / (device_ms / 1000.0), "unit": "rows/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": device_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
        "run": {"launch": "silo_gpu_query_combinations per query (synchronous: the combinations come back to the host)",
                "parallelism": f"interleaved chunk shards (chunk c on rank c % {n_gpus}); (key, count) lists all-gathered (NCCL, {(max_entries + 1) * 16} B per rank), "
                               "summed per key on rank 0" if n_gpus > 1 else "single GPU",
                "combinations": len(rows)},
        "clocks": clocks,
        "e2e": {"value": total_rows * args.steps / (e2e_ms / 1000.0), "unit": "rows/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": 256 + ((max_entries + 1) * 16 if n_gpus > 1 else 0), "d2h_bytes_per_step": 16 * len(rows) * (n_gpus if n_gpus > 1 else 1)},
        "gpu_launches": gpu_launches,
        "roofline": {"bound": "hbm", "achieved": bytes_per_query * args.steps / (device_ms / 1000.0) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bytes_per_query * args.steps / (device_ms / 1000.0) / 1e9 / peak, "traffic": None,
                     "kernel": "whole query (positionCodesKernel + combinationCountKernel + compaction, host synchronisation included)",
                     "algorithmic_bytes_per_launch": bytes_per_query, "kernel_ms": device_ms / args.steps, "peak_source": peak_source},
    }
    if not args.skip_cpu_baseline and n_gpus == 1:
        oracle = oracle_table(total_rows)
        started = time.perf_counter()
        want = oracle.bitmap_aggregation(dimensions, None)
        oracle_seconds = time.perf_counter() - started
        assert rows == want, "the combinations differ from the oracle's"
        line["parity"] = {"oracle": "full size, same table", "combinations_equal": True, "combinations": len(want)}
        line["cpu_baseline"] = {"value": total_rows / oracle_seconds, "unit": "rows/s", "cores": 1, "kind": "port",
                                "sample": f"the whole table ({total_rows} rows), one query, one thread: {oracle_seconds:.1f}s"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if n_gpus > 1:
        dist.destroy_process_group()



def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=20)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", choices=["ours", "reference"], default="ours")
    parser.add_argument("--rows-per-gpu", type=int, default=10_000_000)
    parser.add_argument("--cpu-seconds", type=float, default=12.0)
    parser.add_argument("--reference-step-seconds", type=float, default=2.0)
    parser.add_argument("--skip-cpu-baseline", action="store_true")
    parser.add_argument("--eager", action="store_true", help="launch the timed steps one by one instead of as one CUDA graph")
    parser.add_argument("--workload", choices=["mutations", "nof", "aa", "reads", "cooc"], default="mutations",
                        help="mutations: BASELINE.json configs[1] (the metric's workload, default); nof: configs[2], the NOf / MutationProfile filter; "
                             "aa: configs[3], AminoAcidMutations over 12 genes; reads: configs[2], many_short_read_filters; "
                             "cooc: configs[4], co_occurrence_benchmark over row-partitioned shards")
    parser.add_argument("--reduce", choices=["peer", "nccl"], default="peer",
                        help="N > 1: how the per-rank counts meet -- the library's shard group (peer-memory stores) or NCCL all-reduce")
    parser.add_argument("--force-shard-group", action="store_true",
                        help="N = 1 only, measurement: run the step through a shard group of one rank (what the sharded kernels cost without a peer)")
    parser.add_argument("--traffic-bytes", type=int, default=None,
                        help="dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture")
    args = parser.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.workload == "cooc":
        run_cooc(args)
    elif args.workload == "reads":
        run_reads(args)
    elif args.workload == "aa":
        run_aa(args)
    elif args.workload == "nof":
        if args.impl == "reference":
            run_nof_reference(args)
        else:
            run_nof(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
