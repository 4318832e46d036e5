"""Analytic evaluation of the blocking interleaved (sync-VPP) pipeline schedule + its memory aggregation.

Time: every rank executes the op order produced by ``sim.schedule.interleaved_sync_ops`` (the same generator the
simulator uses); F/B ops advance the rank clock by the chunk's per-micro-batch phase time, a batched p2p group
submits all its sends/recvs at one local timestamp, each send/recv pair completes at
``max(ready_a, ready_b) + max(dur_a, dur_b)`` and the rank resumes at the latest completion of its group.
(reference perf_llm.py:2325-2605)

Memory: walk the local F/B sequence of a stage, adding a chunk's activation cache on F and releasing it on B;
peak = static + max over the walk of (live cache + phase peak inside the chunk).  (reference :1745-1928)
"""
from __future__ import annotations

import heapq
from typing import List

from .utils import convert_final_result_to_human_format

FIRST, MIDDLE, LAST = "first_stage_chunk", "middle_stage_chunk", "last_stage_chunk"


def _stage_key(perf, pp_rank):
    return perf._stage_key_for_pp_rank(pp_rank)


def _plan(perf, pp_rank):
    from ..sim.schedule import InterleavedPlan
    st = perf.strategy
    return InterleavedPlan(st.pp_size, perf._vp_size(), pp_rank, st.micro_batch_num,
                           getattr(st, "microbatch_group_size_per_vp_stage", None))


def compute_interleaved_sync_schedule(perf, return_schedules: bool = False):
    from ..sim.schedule import interleaved_sync_ops
    st = perf.strategy
    pp, vp = st.pp_size, perf._vp_size()
    assert pp > 1 and vp > 1
    if st.pp_comm_async:
        raise RuntimeError("perf interleaved sync schedule does not support async VPP; disable pp_comm_async or "
                           "use simulate() for async VPP runtime behavior")
    queues: List[list] = []
    for r in range(pp):
        names = perf.vpp_stage_chunk_names.get(_stage_key(perf, r), [])
        phases = [perf._compute_single_batch_phase_inputs(n) for n in names]
        p2p = phases[0]["fwd_send"] or phases[0]["fwd_recv"] or phases[0]["bwd_send"] or phases[0]["bwd_recv"]
        q = []
        for op in interleaved_sync_ops(_plan(perf, r)):
            if op[0] in ("F", "B"):
                _, chunk, mb = op
                dur = phases[chunk]["fwd_compute" if op[0] == "F" else "bwd_compute"]
                q.append({"kind": op[0], "mb": mb, "chunk": chunk, "dur": dur, "label": f"{names[chunk]}_model"})
            else:
                subs = []
                for role, tag in op[1]:
                    src, dst = {"send_next": (r, r + 1), "recv_prev": (r - 1, r), "send_prev": (r, r - 1),
                                "recv_next": (r + 1, r)}[role]
                    subs.append({"kind": "send" if role.startswith("send") else "recv", "role": role,
                                 "gid": f"send_recv-{src % pp}-{dst % pp}-{tag}", "dur": p2p, "ready": None,
                                 "done": False})
                q.append({"kind": "comm", "subs": subs, "max_end": 0.0, "wait_gid": None})
        queues.append(q)

    clock = [0.0] * pp
    records: List[List[dict]] = [[] for _ in range(pp)]
    heap, ver, waiting = [], [0] * pp, {}

    def push(r):
        ver[r] += 1
        heapq.heappush(heap, (clock[r], r, ver[r]))

    def rec(r, kind, mb, chunk, label, gid, start, end):
        records[r].append({"kind": kind, "mb": mb, "start": start, "duration": end - start, "end": end,
                           "label": label, "chunk_idx": chunk, "virtual_idx": None, "gid": gid})

    for r in range(pp):
        push(r)
    while any(queues) or waiting:
        if not heap:
            raise RuntimeError(f"interleaved sync scheduler deadlock: {[q[0]['kind'] if q else None for q in queues]}")
        _, r, token = heapq.heappop(heap)
        if token != ver[r] or not queues[r]:
            continue
        op = queues[r][0]
        if op["kind"] in ("F", "B"):
            rec(r, op["kind"], op["mb"], op["chunk"], op["label"], None, clock[r], clock[r] + op["dur"])
            clock[r] += op["dur"]
            queues[r].pop(0)
            push(r)
            continue
        submit = clock[r]
        cursor = submit
        op["max_end"] = max(op["max_end"], submit)
        op["wait_gid"] = None
        for sub in op["subs"]:
            if sub["done"]:
                continue
            if sub["ready"] is None:
                sub["ready"] = submit
            waiter = waiting.get(sub["gid"])
            if waiter is None:
                waiting[sub["gid"]] = (r, op, sub)
                if op["wait_gid"] is None:
                    op["wait_gid"] = sub["gid"]
                continue
            pr, pop_, psub = waiter
            if pr == r:
                if op["wait_gid"] is None:
                    op["wait_gid"] = sub["gid"]
                continue
            if psub["kind"] == sub["kind"]:
                raise RuntimeError(f"same-kind rendezvous for gid {sub['gid']}: {sub['kind']}")
            waiting.pop(sub["gid"], None)
            end = max(sub["ready"], psub["ready"]) + max(sub["dur"], psub["dur"])
            rec(r, sub["kind"], None, None, sub["role"], sub["gid"], sub["ready"], end)
            rec(pr, psub["kind"], None, None, psub["role"], psub["gid"], psub["ready"], end)
            sub["done"] = psub["done"] = True
            cursor = max(cursor, end)
            op["max_end"] = max(op["max_end"], end)
            clock[pr] = max(clock[pr], end)
            pop_["max_end"] = max(pop_["max_end"], end)
            if pop_["wait_gid"] == sub["gid"] and queues[pr] and queues[pr][0] is pop_:
                push(pr)
        clock[r] = max(clock[r], cursor)
        if all(s["done"] for s in op["subs"]):
            clock[r] = max(clock[r], op["max_end"])
            queues[r].pop(0)
            push(r)
    total = max(clock) if clock else 0.0
    return (total, records) if return_schedules else total


# ------------------------------------------------------------------------------------------------
def _chunk_memory_profile(perf, name: str) -> dict:
    pk = perf.pp_state_peak_point.get(name)
    if pk is None:
        chunk = perf.model_chunk_dict.get(name) or perf.vpp_chunk_dict[name]
        pk = chunk.compute_activations()
        perf.pp_state_peak_point[name] = pk
    cache = pk.activation_mem_cache
    window = max(pk.bwd_peak_mem, pk.recomp_fwd_peak_mem, pk.recomp_bwd_peak_mem)
    if window == pk.recomp_fwd_peak_mem:
        bpath, bstage = pk.recomp_fwd_peak_path, "recompute_forward"
    elif window == pk.recomp_bwd_peak_mem:
        bpath, bstage = pk.recomp_bwd_peak_path, "recompute_backward"
    else:
        bpath, bstage = pk.bwd_peak_path, "backward"
    return {"cache": cache, "fwd_peak": pk.fwd_peak_mem, "bwd_peak": max(0.0, window - cache),
            "fwd_path": pk.fwd_peak_path, "bwd_path": bpath, "bwd_stage": bstage}


def local_phase_sequence(perf, pp_rank: int) -> list:
    plan = _plan(perf, pp_rank)
    seq = [("fwd",) + plan.fwd_ref(k)[:2] for k in range(plan.warmup)]
    for k in range(plan.remaining):
        seq.append(("fwd",) + plan.fwd_ref(k + plan.warmup)[:2])
        seq.append(("bwd",) + plan.bwd_ref(k)[:2])
    seq += [("bwd",) + plan.bwd_ref(k)[:2] for k in range(plan.remaining, plan.total)]
    return seq          # items: (phase, microbatch, chunk_idx)


def analysis_sync_vpp_stage_mem(perf, pp_rank: int) -> dict:
    st = perf.strategy
    key = _stage_key(perf, pp_rank)
    names = list(perf.vpp_stage_chunk_names.get(key, []))
    if not names:
        return {}
    dense = dict(weight_bytes=0.0, grad_bytes=0.0, state_bytes=0.0)
    moe = dict(weight_bytes=0.0, grad_bytes=0.0, state_bytes=0.0)
    shapes = set()
    for n in names:
        mi = perf.vpp_chunk_dict[n].get_model_info()
        dense["weight_bytes"] += mi.dense_weight_bytes; dense["grad_bytes"] += mi.dense_grad_bytes
        dense["state_bytes"] += mi.dense_state_bytes
        moe["weight_bytes"] += mi.moe_weight_bytes; moe["grad_bytes"] += mi.moe_grad_bytes
        moe["state_bytes"] += mi.moe_state_bytes
        shapes |= mi.te_dummy_wgrad_shapes
    dummy_bytes = sum(r * c * e for r, c, e in shapes)
    res = {"micro_batch_num": st.micro_batch_num, "micro_batch_size": st.micro_batch_size,
           "parallel_config": {"parallelism": st.parallelism, "fp8": st.fp8, "recompute_status": {
               "layer_num": perf.model_config.layer_num,
               "actual_layer_num": sum(perf.vpp_chunk_dict[n].layer_num for n in names),
               "recompute_layer": st.recompute_layer_num,
               "recompute_recompute_granularity": st.recompute_granularity}},
           "memory_schedule": "sync_vpp_schedule", "stage_type": key, "stage_rank": pp_rank}
    model_mem = sum(dense.values()) + sum(moe.values()) + dummy_bytes
    res["model_mem"] = model_mem
    res["model_mem_detail"] = dict(dense=dict(all_mem=sum(dense.values()), detail=dense),
                                   moe=dict(all_mem=sum(moe.values()), detail=moe),
                                   te_dummy_wgrad=dict(all_mem=dummy_bytes, detail=dict(
                                       dummy_wgrad_bytes=dummy_bytes, shape_count=len(shapes),
                                       shapes=[list(x) for x in sorted(shapes)])))
    prof = {n: _chunk_memory_profile(perf, n) for n in names}
    gbs = sorted({p["cache"] / 1024 ** 3 for p in prof.values()})
    res["fwd_activation_cache_per_micro_batch"] = (f"{gbs[0]:.4f} GB" if len(gbs) == 1
                                                   else f"{gbs[0]:.4f} ~ {gbs[-1]:.4f} GB")
    live, entries, max_entries, peak, peak_path, peak_stage = 0.0, 0, 0, 0.0, "", ""
    for phase, mb, chunk in local_phase_sequence(perf, pp_rank):
        p = prof[names[chunk]]
        if phase == "fwd":
            in_chunk, path, stage, delta = p["fwd_peak"], p["fwd_path"], "forward", p["cache"]
            if delta > 0:
                entries += 1
        else:
            in_chunk, path, stage, delta = p["bwd_peak"], p["bwd_path"], p["bwd_stage"], -p["cache"]
            if delta < 0 and p["cache"] > 0:
                entries -= 1
        if live + in_chunk >= peak:
            peak = live + in_chunk
            peak_path, peak_stage = f"{names[chunk]}[mb{mb},chunk{chunk}]: {path}", stage
        live += delta
        max_entries = max(max_entries, entries)
    assert abs(live) < 1e-6 and entries == 0, f"sync VPP memory aggregation should end with zero live cache, got {live}"
    res["cached_micro_batch_num"] = max_entries
    res["peak_activation_mem_in_1F1B"] = peak
    res["peak_mem"] = model_mem + peak
    res["peak_mem_with_reserved"] = res["peak_mem"] / st.mem_factor
    res["memory_reserved_ratio"] = str(st.mem_factor)
    res["peak_path"] = f"{peak_path}, stage=[{peak_stage}]"
    raw = {"model_mem": model_mem, "peak_mem": res["peak_mem"], "peak_mem_with_reserved": res["peak_mem_with_reserved"]}
    convert_final_result_to_human_format(res)
    res["raw"] = raw
    return res
