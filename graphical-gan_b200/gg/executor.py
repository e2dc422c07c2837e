"""Plan compiler / executor: fetch set -> static launch list over the C-ABI -> CUDA graph.

`Session.run(fetches, feed_dict)` (gmgan_inference_cifar10.py:485-494) maps to `Runtime.run`: the fetch set is
compiled ONCE into a Plan (topologically sorted, pruned to what the fetches need — exactly TF's pruning, e.g.
`rec_x` is never computed in local_ep mode), all buffers are allocated up front (shapes are static), and the
launch list is captured into a CUDA graph so one iteration costs one graph launch instead of a few hundred
kernel launches.  Torch tensors are device-memory containers only; all arithmetic runs in libgg_b200.so.
"""
import ctypes as C
import os
import struct

import numpy as np

from . import cabi
from . import dist as ggdist
from .graph import Tensor, Operation, float32, int32, prod
from .ops import toposort

ALIAS_OPS = ("reshape", "stop_gradient", "aux")


def _torch():
    import torch
    return torch


class Runtime(object):
    """Process-global device state: parameter buffers, optimiser slots, constants, feeds, RNG counter."""

    def __init__(self):
        self.params = {}       # node.id -> device tensor
        self.param_nodes = {}  # node.id -> node
        self.consts = {}
        self.feeds = {}        # node.id -> (device tensor, pinned host tensor)
        self.stage = {}        # key -> ring of pinned staging slots [tensor, event of the last copy, pending Deferred]
        self.slots = {}        # (optimizer id, param id) -> (m, v)
        self.opt_state = {}    # optimizer id -> device state tensor {b1^t, b2^t, t}
        self.plans = {}
        self.pending_restore = {}   # optimiser state read by Saver.restore before the plan that owns it exists
        self.seed = 1234
        self._tick = None
        self._zero = None
        self.use_cuda_graph = os.environ.get("GG_CUDA_GRAPH", "1") != "0"
        self.sync_bn = os.environ.get("GG_SYNC_BN", "1") != "0"
        self.device = None

    # ---- device helpers -----------------------------------------------------------------------
    def dev(self):
        torch = _torch()
        if self.device is None:
            if not torch.cuda.is_available():
                raise cabi.GGError("no CUDA device: the graphical-gan_b200 hot path runs on sm_100a only (no CPU fallback)")
            self.device = torch.device("cuda", torch.cuda.current_device())
        return self.device

    def empty(self, shape, dtype=float32):
        torch = _torch()
        n = max(prod(shape), 1)
        t = torch.empty(n, dtype=torch.float32 if dtype == float32 else torch.int32, device=self.dev())
        return t

    def tick(self):
        torch = _torch()
        if self._tick is None:
            self._tick = torch.zeros(1, dtype=torch.int64, device=self.dev())
        return self._tick

    def zero(self):
        torch = _torch()
        if self._zero is None:
            self._zero = torch.zeros(4, dtype=torch.float32, device=self.dev())
        return self._zero

    def param_buffer(self, node):
        torch = _torch()
        if node.id not in self.params:
            t = torch.from_numpy(np.ascontiguousarray(node.attrs["init"]).reshape(-1)).to(self.dev())
            if t.numel() == 0:
                t = torch.zeros(1, device=self.dev())
            self.params[node.id] = t
            self.param_nodes[node.id] = node
        return self.params[node.id]

    def const_buffer(self, node):
        torch = _torch()
        if node.id not in self.consts:
            arr = np.ascontiguousarray(node.attrs["value"]).reshape(-1)
            if arr.size == 0:
                arr = np.zeros(1, arr.dtype)
            self.consts[node.id] = torch.from_numpy(arr.copy()).to(self.dev())
        return self.consts[node.id]

    def feed_buffer_u8(self, node):
        """uint8 staging pair (device, pinned host) of an int32 placeholder fed with a uint8 array: one byte per element
        crosses PCIe and gg_widen_u8_i32 fills the int32 feed buffer on the device (bit exact)"""
        torch = _torch()
        key = ("u8", node.id)
        if key not in self.feeds:
            n = max(node.size, 1)
            self.feeds[key] = (torch.zeros(n, dtype=torch.uint8, device=self.dev()), torch.zeros(n, dtype=torch.uint8).pin_memory())
        return self.feeds[key]

    def stage_slot(self, key, n, dtype):
        """next pinned staging slot of a ring of FETCH_RING (host->device feeds, device->host fetches): with deferred fetches
        the host runs ahead of the device, so a slot is reused only after the copy that last used it has completed"""
        torch = _torch()
        ring = self.stage.get(key)
        if ring is None:
            ring = self.stage[key] = {"next": 0, "slots": [[torch.zeros(max(n, 1), dtype=dtype).pin_memory(), None, None]
                                                             for _ in range(FETCH_RING)]}
        slot = ring["slots"][ring["next"]]
        ring["next"] = (ring["next"] + 1) % FETCH_RING
        if slot[2] is not None:
            slot[2].result()                 # a Deferred nobody has read yet still points at this slot: settle it first
            slot[2] = None
        if slot[1] is not None:
            slot[1].synchronize()
        return slot

    def feed_buffer(self, node):
        torch = _torch()
        if node.id not in self.feeds:
            dt = torch.float32 if node.dtype == float32 else torch.int32
            d = torch.zeros(max(node.size, 1), dtype=dt, device=self.dev())
            h = torch.zeros(max(node.size, 1), dtype=dt).pin_memory()
            self.feeds[node.id] = (d, h)
        return self.feeds[node.id]

    def rng_seed(self):
        """Philox key of the in-graph random ops: the graph seed with the data-parallel rank in the high word, so that every
        rank draws its OWN slice of the global noise batch (p_z, Gumbel uniforms, dequantisation noise) — with identical
        keys the global fake batch would be `world` copies of one local batch."""
        return (int(self.seed) + (ggdist.rank() << 32)) & 0xFFFFFFFFFFFFFFFF

    # ---- public -------------------------------------------------------------------------------
    def get_param(self, node):
        return self.param_buffer(node).cpu().numpy().reshape(node.shape)

    def set_param(self, node, value):
        torch = _torch()
        self.param_buffer(node).copy_(torch.from_numpy(np.ascontiguousarray(value, dtype=np.float32).reshape(-1)))

    def run(self, fetches, feed_dict=None, to_host=True, deferred=False):
        """to_host=False: enqueue the step and return its Plan without any device->host read (kernel-only timing).
        deferred=True: return Deferred values (device->host copies enqueued, not awaited)."""
        single = isinstance(fetches, (Tensor, Operation)) or fetches is None
        flist = [fetches] if single else list(fetches)
        flat = []

        def flatten(f):
            if isinstance(f, (list, tuple)):
                return [flatten(x) for x in f]
            flat.append(f)
            return len(flat) - 1
        structure = [flatten(f) for f in flist]
        feed_dict = feed_dict or {}
        key = (tuple(f.id for f in flat), tuple(sorted(k.id for k in feed_dict)))
        plan = self.plans.get(key)
        if plan is None:
            plan = Plan(self, flat, list(feed_dict.keys()))
            self.plans[key] = plan
        results = plan.run(feed_dict, to_host=to_host, deferred=deferred)
        if not to_host:
            return plan

        def rebuild(s):
            if isinstance(s, list):
                return [rebuild(x) for x in s]
            return results[s]
        out = [rebuild(s) for s in structure]
        return out[0] if single else out


FETCH_RING = 4      # pinned host slots per fetched tensor / per fed placeholder: how far the host may run ahead of the device


class Deferred(object):
    """Value of a fetched tensor whose device->host copy is enqueued but not yet awaited (Session(deferred_fetches=True)).

    The reference loop only logs the costs `session.run` returns (gmgan_inference_cifar10.py:483-494), so nothing on the
    host has to wait for a step before it feeds the next one.  A Deferred converts to its numpy value on first use
    (float(), np.asarray(), arithmetic, comparison, formatting, any numpy attribute): that is the moment the host waits for the
    copy's CUDA event.  The copy itself is enqueued by EVERY run, into pinned memory."""
    __slots__ = ("_h", "_ev", "_shape", "_val")

    def __init__(self, h, ev, shape):
        self._h, self._ev, self._shape, self._val = h, ev, tuple(shape), None

    def result(self):
        if self._val is None:
            self._ev.synchronize()
            arr = self._h.numpy().reshape(self._shape).copy()
            self._val = arr[()] if len(self._shape) == 0 else arr
            self._h = self._ev = None
        return self._val

    def done(self):
        return self._val is not None or self._ev.query()

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self.result())
        return a.astype(dtype) if dtype is not None and a.dtype != dtype else a

    def __getattr__(self, name):            # .shape, .dtype, .item(), .mean(), ... of the numpy value
        return getattr(self.result(), name)

    def __float__(self): return float(self.result())
    def __int__(self): return int(self.result())
    def __bool__(self): return bool(self.result())
    def __len__(self): return len(self.result())
    def __getitem__(self, i): return self.result()[i]
    def __iter__(self): return iter(self.result())
    def __repr__(self): return repr(self.result())
    def __str__(self): return str(self.result())
    def __format__(self, spec): return format(self.result(), spec)
    def __neg__(self): return -self.result()
    def __abs__(self): return abs(self.result())
    def __add__(self, o): return self.result() + o
    def __radd__(self, o): return o + self.result()
    def __sub__(self, o): return self.result() - o
    def __rsub__(self, o): return o - self.result()
    def __mul__(self, o): return self.result() * o
    def __rmul__(self, o): return o * self.result()
    def __truediv__(self, o): return self.result() / o
    def __rtruediv__(self, o): return o / self.result()
    def __pow__(self, o): return self.result() ** o
    def __lt__(self, o): return self.result() < o
    def __le__(self, o): return self.result() <= o
    def __gt__(self, o): return self.result() > o
    def __ge__(self, o): return self.result() >= o
    def __eq__(self, o): return self.result() == o
    def __ne__(self, o): return self.result() != o
    __hash__ = None


RT = Runtime()
ALL_PLANS = []      # every Plan ever built: gg.dist.shutdown() drops their captured graphs before NCCL goes away


def reset_runtime():
    global RT
    RT.__init__()


class Plan(object):
    def __init__(self, rt, fetches, fed):
        torch = _torch()
        self.rt = rt
        self.fetches = fetches
        self.fed = {f.id: f for f in fed}
        self.buf = {}        # node.id -> device tensor (flat)
        self.extra = {}      # (node.id, k) -> device tensor
        self.steps = []      # callables f(stream_ptr)
        self.keep = []       # keep ctypes arrays / tables alive
        self.has_random = False
        self.n_launch_est = 0
        roots = []
        for f in fetches:
            if isinstance(f, Operation):
                roots.extend(self._op_roots(f))
            elif isinstance(f, Tensor):
                roots.append(f)
        self.order = self._toposort(roots)
        # split-K of the tensor-core kernels may use every SM (GG_TC_MAX_CTAS caps it); with the 3-stage ring two CTAs share an
        # SM, so two launches on different streams still overlap.  Workspace sizes depend on both: fixed before emission
        self.n_streams = int(os.environ.get("GG_STREAMS", "8")) if rt.use_cuda_graph else 1
        cabi.call("gg_set_tc_max_ctas", int(os.environ.get("GG_TC_MAX_CTAS", "148")))
        cabi.call("gg_set_pdl", 1 if (os.environ.get("GG_PDL", "0") == "1" and rt.use_cuda_graph) else 0)
        cabi.call("gg_set_tc_stages", int(os.environ.get("GG_TC_STAGES", "3" if self.n_streams > 1 else "0")))
        # scheduling metadata: one group per node that launches kernels; `owner` resolves views (reshape / aux / fed)
        # to the node whose kernels produce the storage
        self.groups = []     # dicts: start, end, reads (owner ids), writes (owner id), barrier
        self.owner = {}
        self.read_override = {}   # node id -> the tensors its launch really reads (when a peephole bypasses an input node)
        self.self_grouped = set()
        # the in-graph random ops key their Philox streams with a per-run counter.  It is advanced AFTER the last draw of a run
        # (GG_TICK_FIRST=1: before the first one, as in round 1): at the head of the step the one-thread tick launch sat on the
        # critical chain in front of the prior sample; at the tail it depends only on the draws and overlaps the backward pass
        tick_first = os.environ.get("GG_TICK_FIRST", "0") == "1"
        if any(n.op == "random" and n.id not in self.fed for n in self.order):
            self.has_random = True
            if tick_first:
                self._emit_tick(set())
        self._plan_inplace_concats()
        self._plan_grad_buckets(fetches)
        self._plan_actgrad_fusion(fetches)
        self._plan_transpose_fusion(fetches)
        self._plan_bn_stats()
        self._plan_ew_fusion(fetches)
        self._resort()
        for node in self.order:
            cl = self.ew_cluster_of.get(node.id)
            if cl is not None:
                if not cl.emitted:
                    self._emit_ew_cluster(cl)
                continue
            s0 = len(self.steps)
            self._emit(node)
            self._note_group(node, s0)
        if self.has_random and not tick_first:
            self._emit_tick(set(n.id for n in self.order if n.op == "random" and n.id not in self.fed))
        for f in fetches:
            if isinstance(f, Operation):
                self._emit_operation(f)
        self.graph = None
        ALL_PLANS.append(self)
        self.trace, self.trace_t0 = [], None
        self.kernel_launches = 0   # libgg_b200 kernels per run (counted at capture / eager launch)
        self.runs = 0

    def _emit_tick(self, reads):
        tick = self.rt.tick()
        s0 = len(self.steps)
        self.steps.append(lambda st, tick=tick: cabi.call("gg_rng_tick", tick.data_ptr(), st))
        self.groups.append(dict(start=s0, end=s0 + 1, reads=set(reads), writes="tick", barrier=False, collective=False, node=None,
                                part=(0, 1), ordered=False))

    def _owners(self, t):
        return self.owner.get(t.id, frozenset([t.id]))

    def _note_group(self, node, s0):
        if node.id in self.self_grouped:          # the launcher registered its own groups (a concat: one per piece)
            return
        if len(self.steps) == s0:        # no kernels: a leaf, or a view (reshape, axis-0 slice, in-place concat, ...)
            if node.inputs and node.id not in self.fed:
                own = frozenset()
                for i in node.inputs:
                    own |= self._owners(i)
                self.owner[node.id] = own
            else:
                self.owner[node.id] = frozenset([node.id])
            return
        self.owner[node.id] = frozenset([node.id])
        reads = set()
        for i in self.read_override.get(node.id, node.inputs):
            reads |= self._owners(i)
        if node.id in getattr(self, "fuse_mask", {}):
            reads |= self._owners(self.fuse_mask[node.id][0])
        if node.op == "random":
            reads.add("tick")
        self._add_groups(s0, len(self.steps), reads, node.id, barrier=False, node=node)

    def _add_groups(self, s0, end, reads, writes, barrier, node=None):
        """one scheduling group per maximal run of kernel steps; a collective step inside a node (SyncBN statistics, the
        gradient bucket all-reduce) becomes its own group so the launch list can be cut there"""
        cuts = [i for i in range(s0, end) if getattr(self.steps[i], "is_collective", False)]
        ranges, a = [], s0
        for c in cuts:
            if c > a:
                ranges.append((a, c, False))
            ranges.append((c, c + 1, True))
            a = c + 1
        if end > a:
            ranges.append((a, end, False))
        prev = None
        for k, (a, b, coll) in enumerate(ranges):
            last = k == len(ranges) - 1
            w = writes if last else "%s#%d" % (writes, k)
            r = set(reads) if prev is None else {prev}
            # peer-memory all-reduce kernels share one exchange buffer and an epoch counter per rank: they must run one at
            # a time and in the same order on every rank, so the scheduler chains the groups that contain one
            # ordering classes: NCCL collectives among themselves (one communicator), peer-memory exchange kernels among
            # themselves (same order, one at a time on every rank); the two classes are independent of each other so a
            # gradient bucket's all-reduce can start while later SyncBN layers of the backward pass are still running
            ordered = "nccl" if coll else ("peer" if any(getattr(self.steps[i], "is_small_allreduce", False) for i in range(a, b)) else None)
            self.groups.append(dict(start=a, end=b, reads=r, writes=w, barrier=barrier and prev is None, collective=coll,
                                    node=node, part=(k, len(ranges)), ordered=ordered))
            prev = w

    # ---- graph walking ---------------------------------------------------------------------
    def _op_roots(self, op):
        roots = [d for d in op.deps if d is not None]
        for sub in op.attrs.get("ops", ()):
            roots.extend(self._op_roots(sub))
        return roots

    def _toposort(self, roots):
        # like ops.toposort but does not look behind fed tensors (TF lets you feed any tensor)
        order, seen = [], set()
        stack = [(r, False) for r in reversed(roots)]
        while stack:
            node, done = stack.pop()
            if done:
                order.append(node)
                continue
            if node.id in seen:
                continue
            seen.add(node.id)
            stack.append((node, True))
            if node.id in self.fed:
                continue
            for inp in reversed(node.inputs):
                if inp.id not in seen:
                    stack.append((inp, False))
        return order

    # ---- emission ---------------------------------------------------------------------------
    def _alloc(self, node):
        if node.id in self.placed_flat:   # a parameter gradient produced straight into its optimiser's all-reduce bucket
            oid, off = self.placed_flat[node.id]
            t = self.flat[oid][off:off + max(node.size, 1)]
            self.buf[node.id] = t
            return t
        if node.id in self.placed:       # this node's output lives inside the buffer of an axis-0 concat (zero-copy concat)
            cid, off = self.placed[node.id]
            t = self._concat_storage(cid)[off:off + max(node.size, 1)]
        else:
            t = self.rt.empty(node.shape, node.dtype)
        self.buf[node.id] = t
        return t

    # ---- zero-copy row concatenation / row slices --------------------------------------------------------------------
    # The sibling-batching rewrite (gg/rewrite.py) joins the inputs of the two discriminator towers along axis 0 and
    # splits the batched result again.  Both are free here: the producers of a concat's pieces write straight into the
    # concat's buffer, and an axis-0 slice of a contiguous tensor is a pointer offset.
    OWN_OUTPUT_OPS = ("unary", "binary", "broadcast", "add_n", "cast", "reduce", "softmax", "softmax_grad", "one_hot", "random",
                      "transpose", "pad", "tile", "matmul", "conv", "bn", "bn_grad")

    @staticmethod
    def _is_row_concat(node):
        return node.op == "concat" and node.dtype == float32 and prod(node.shape[:node.attrs["axis"]]) == 1

    @staticmethod
    def _slice_is_view(node):
        a = node.attrs
        shp = node.inputs[0].shape
        inner = prod(shp[a["axis"] + 1:])
        return prod(shp[:a["axis"]]) == 1 and (a["start"] * inner) % 64 == 0 and node.dtype == float32

    def _plan_inplace_concats(self):
        self.placed, self.inplace_concat, self._concat_bufs = {}, set(), {}
        if os.environ.get("GG_INPLACE_CONCAT", "1") == "0":
            return
        for node in self.order:
            if not self._is_row_concat(node) or node.id in self.fed:
                continue
            off, plan, ok = 0, [], True
            for inp in node.inputs:
                if inp.op not in self.OWN_OUTPUT_OPS or inp.id in self.fed or inp.id in self.placed or inp.dtype != float32 \
                        or off % 64 != 0 or any(inp is q for q, _ in plan) or (inp.op == "slice" and self._slice_is_view(inp)):
                    ok = False
                    break
                plan.append((inp, off))
                off += inp.size
            if ok:
                for inp, o in plan:
                    self.placed[inp.id] = (node.id, o)
                self.inplace_concat.add(node.id)

    # ---- data-parallel gradient buckets --------------------------------------------------------------------------------
    # One flat fp32 buffer per optimiser (parameters replicated, gradients summed once per step).  Round 1
    # packed the gradients into it with an extra kernel after the whole backward pass and all-reduced it in one exposed,
    # eager NCCL call between two graph segments.  Now (a) every gradient whose only consumer is the optimiser is PRODUCED
    # inside the buffer (the wgrad / bias-sum kernel writes there: no pack pass), (b) the buffer is laid out in the order the
    # gradients become available and cut into GG_DP_BUCKETS ranges, each all-reduced as soon as its last producer has run —
    # on a stream of its own inside the step's CUDA graph, beside the rest of the backward pass.
    def _optimizer_ops(self, fetches):
        out = []

        def walk(op):
            if op.kind == "group":
                for sub in op.attrs["ops"]:
                    walk(sub)
            elif op.kind in ("adam", "rmsprop"):
                out.append(op)
        for f in fetches:
            if isinstance(f, Operation):
                walk(f)
        return out

    def _plan_grad_buckets(self, fetches):
        torch = _torch()
        self.placed_flat, self.flat, self.bucket_plan = {}, {}, {}
        world = ggdist.world_size()
        # readiness of a node = length of the longest dependency chain below it (tensor-core launches weigh ~6 glue
        # launches): the toposort position is useless here — a depth-first order emits the deepest gradient's whole chain first
        pos = {}
        for n in self.order:
            w = 6 if n.op in ("conv", "matmul") else (0 if n.op in ALIAS_OPS or not n.inputs else 1)
            pos[n.id] = w + max([pos.get(i.id, 0) for i in n.inputs] or [0])
        self.ready_pos = pos
        if world <= 1:
            return
        uses = {}
        for n in self.order:
            if n.id in self.fed:
                continue
            for i in n.inputs:
                uses[i.id] = uses.get(i.id, 0) + 1
        opts = self._optimizer_ops(fetches)
        for f in fetches:
            if isinstance(f, Tensor):
                uses[f.id] = uses.get(f.id, 0) + 1
        for op in opts:
            for d in op.deps:
                if d is not None:
                    uses[d.id] = uses.get(d.id, 0) + 1
        direct = os.environ.get("GG_DP_DIRECT", "1") != "0"
        n_buckets = max(1, int(os.environ.get("GG_DP_BUCKETS", "4")))
        for op in opts:
            entries = []
            for v, g in zip(op.attrs["vars"], op.deps):
                if g is None:
                    continue
                own, ok = g, uses.get(g.id, 0) == 1
                while own.op in ("reshape", "stop_gradient") and own.id not in self.fed:
                    own = own.inputs[0]
                    ok = ok and uses.get(own.id, 0) == 1
                ok = ok and direct and own.op in self.OWN_OUTPUT_OPS and own.id not in self.fed and own.id not in self.placed \
                    and own.id not in self.placed_flat and own.dtype == float32 and own.size == v.size
                entries.append(dict(var=v, grad=g, own=own, direct=ok, ready=pos.get(own.id, 0)))
            entries.sort(key=lambda e: e["ready"])
            off = 0
            for e in entries:
                e["off"] = off
                off += (e["var"].size + 63) & ~63                 # 256-byte aligned slots
            total = max(off, 64)
            self.flat[op.attrs["opt_id"]] = torch.zeros(total, dtype=torch.float32, device=self.rt.dev())
            for e in entries:
                if e["direct"]:
                    self.placed_flat[e["own"].id] = (op.attrs["opt_id"], e["off"])
            # cut into buckets of roughly equal bytes in readiness order; the gradients that become ready LAST (split update,
            # _late_vars) form a small bucket of their own, so that the exchange exposed after the last backward kernel is a
            # few KB instead of a quarter of the model
            # GG_DP_LATE_BUCKET=1: the small last bucket without the split update
            late = self._late_vars([(e["var"], e["own"]) for e in entries], force=os.environ.get("GG_DP_LATE_BUCKET", "0") == "1")
            main = [e for e in entries if e["var"].id not in late]
            tail = [e for e in entries if e["var"].id in late]
            buckets, cur, acc, per = [], [], 0, total / float(n_buckets)
            for e in main:
                cur.append(e)
                acc += (e["var"].size + 63) & ~63
                if acc >= per * (len(buckets) + 1) and len(buckets) < n_buckets - 1:
                    buckets.append(cur)
                    cur = []
            if cur:
                buckets.append(cur)
            if tail:
                buckets.append(tail)
            self.bucket_plan[op.attrs["opt_id"]] = (entries, buckets, total)

    def _plan_actgrad_fusion(self, fetches):
        """dgrad -> activation-gradient pairs that run as ONE launch: `leaky_grad(y, conv_dgrad(dy, w))` (autodiff of
        LeakyReLU(Conv2D(.)) / ReLU(Deconv2D(.)) stacks) becomes gg_conv2d_dgrad_actgrad, whose write-out multiplies by act'(y).
        Four such pairs sit on the critical chain of each training step (3-5.5 us + a dependent launch each).  Conditions: the
        dgrad runs on the tensor-core kernels, has no bias / activation of its own, is consumed only by the gradient node,
        neither output is placed inside another buffer, and y is computed before the dgrad in the plan's order."""
        self.fuse_mask, self.fused_alias = {}, {}
        if os.environ.get("GG_FUSE_ACTGRAD", "1") == "0":
            return
        pos = {n.id: i for i, n in enumerate(self.order)}
        uses = {}
        for n in self.order:
            if n.id in self.fed:
                continue
            for i in n.inputs:
                uses[i.id] = uses.get(i.id, 0) + 1
        for f in fetches:
            if isinstance(f, Tensor):
                uses[f.id] = uses.get(f.id, 0) + 1
            else:
                for d in self._op_roots(f):
                    uses[d.id] = uses.get(d.id, 0) + 1
        placed = set(getattr(self, "placed", {})) | set(getattr(self, "placed_flat", {}))
        for n in self.order:
            if n.op != "binary" or n.attrs.get("fn") not in ("leaky_grad", "relu_grad") or n.id in self.fed:
                continue
            y, g = n.inputs
            dense = g.op == "matmul" and not g.attrs["ta"] and g.attrs["tb"]       # dx = dy W^T of a Linear layer (linear.py:133)
            if not dense and (g.op != "conv" or g.attrs["mode"] != "dgrad"):
                continue
            if g.attrs["act"] is not None or len(g.inputs) != 2:
                continue
            if g.id in self.fed or y.id not in pos or uses.get(g.id, 0) != 1 or g.id in placed or n.id in placed:
                continue
            if tuple(y.shape) != tuple(g.shape) or tuple(n.shape) != tuple(g.shape) or pos[y.id] > pos[g.id]:
                continue
            if dense:
                # a dense layer is a 1x1 convolution on a 1x1 image (gg_gemm maps dy W^T onto the same dgrad launch)
                M, N = g.shape
                K = g.inputs[0].shape[1]
                if os.environ.get("GG_FUSE_ACTGRAD_DENSE", "1") == "0" or cabi.lib.gg_conv2d_tc_supported(1, M, 1, 1, N, K, 1, 1, 1, 1) != 1:
                    continue
                self.fuse_mask[g.id] = (y, n.attrs["fn"][:-5], float(n.attrs["alpha"]))
                self.fused_alias[n.id] = g.id
                continue
            a = g.attrs
            if cabi.lib.gg_conv2d_tc_supported(1, a["B"], a["H"], a["W"], a["Ci"], a["Co"], a["k"], a["stride"], a["Ho"], a["Wo"]) != 1:
                continue
            self.fuse_mask[g.id] = (y, n.attrs["fn"][:-5], float(n.attrs["alpha"]))
            self.fused_alias[n.id] = g.id

    @staticmethod
    def _tc_geometry(node):
        """(mode, geometry tuple) of a conv / dense node as the tensor-core kernels see it, or None (wgrad-type products)"""
        if node.op == "conv":
            g = node.attrs
            if g["mode"] not in ("fwd", "dgrad"):
                return None
            return (0 if g["mode"] == "fwd" else 1,
                    (g["B"], g["H"], g["W"], g["Ci"], g["Co"], g["k"], g["stride"], g["pad_t"], g["pad_l"], g["Ho"], g["Wo"]))
        if node.op == "matmul" and not node.attrs["ta"]:
            M, N = node.shape
            K = node.inputs[0].shape[1]
            # a dense layer is a 1x1 convolution on a 1x1 image: y = x W is its fwd (Ci = K, Co = N), dx = dy W^T its dgrad
            return (1, (M, 1, 1, N, K, 1, 1, 0, 0, 1, 1)) if node.attrs["tb"] else (0, (M, 1, 1, K, N, 1, 1, 0, 0, 1, 1))
        return None

    def _plan_bn_stats(self):
        """`Batchnorm` always normalises the output of a Conv2D / Deconv2D / Linear (tflib/ops/batchnorm.py:29-30 after
        conv2d.py:106-120).  When that producer runs on the tensor-core kernels its epilogue also writes the per-m-tile column
        sums and sums of squares (gg_conv2d_bnstats), and the batch norm becomes ONE element-wise pass (gg_bn_apply folds the
        tile rows in double precision and normalises): the statistics pass over the activation — and the cluster rendezvous of
        the one-launch batch-norm kernel — disappear from the step's critical chain."""
        self.bn_stats, self.bn_from_stats = {}, {}
        if os.environ.get("GG_BN_CONV_STATS", "1") == "0" or (ggdist.world_size() > 1 and self.rt.sync_bn):
            return
        placed = set(self.placed) | set(self.placed_flat)
        for n in self.order:
            if n.op != "bn" or n.id in self.fed:
                continue
            x = n.inputs[0]
            prod_node = x
            if x.op == "reshape" and x.id not in self.fed and x.shape[-1] == x.inputs[0].shape[-1]:
                prod_node = x.inputs[0]
            if prod_node.op not in ("conv", "matmul") or prod_node.id in self.fed or prod_node.id in self.fuse_mask \
                    or prod_node.id in self.bn_stats or prod_node.id in placed or prod_node.dtype != float32:
                continue
            geo = self._tc_geometry(prod_node)
            Cc = n.shape[-1]
            if geo is None or prod_node.shape[-1] != Cc or prod_node.size != n.size:
                continue
            tiles = cabi.lib.gg_conv2d_stats_tiles(geo[0], *geo[1])
            if tiles <= 0:
                continue
            self.bn_stats[prod_node.id] = dict(mode=geo[0], geo=geo[1], tiles=tiles, C=Cc, buf=None)
            self.bn_from_stats[n.id] = prod_node.id

    def _use_counts(self, fetches):
        uses = {}
        for n in self.order:
            if n.id in self.fed:
                continue
            for i in n.inputs:
                uses[i.id] = uses.get(i.id, 0) + 1
        for f in fetches:
            if isinstance(f, Tensor):
                uses[f.id] = uses.get(f.id, 0) + 1
            else:
                for d in self._op_roots(f):
                    uses[d.id] = uses.get(d.id, 0) + 1
        for _gid, (y, _a, _al) in self.fuse_mask.items():
            uses[y.id] = uses.get(y.id, 0) + 1
        return uses

    def _is_ancestor(self, anc, node):
        """does `node` depend on `anc`? (walks inputs; fed nodes cut the walk like they cut the plan)"""
        seen, stack = set(), [node]
        while stack:
            n = stack.pop()
            if n is anc:
                return True
            if n.id in seen or n.id in self.fed:
                continue
            seen.add(n.id)
            stack.extend(n.inputs)
        return False

    @staticmethod
    def _b2d_form(node):
        """(Bt, R, C) when the transpose is a batched 2-D one ([Bt, R, C] -> [Bt, C, R]), else None"""
        shp, perm = list(node.inputs[0].shape), list(node.attrs["perm"])
        nd = len(shp)
        if nd >= 3 and perm[0] == 0 and perm[1:] == list(range(2, nd)) + [1]:
            return shp[0], shp[1], prod(shp[2:])
        if nd >= 3 and perm[0] == 0 and perm[1:] == [nd - 1] + list(range(1, nd - 1)):
            return shp[0], prod(shp[1:nd - 1]), shp[nd - 1]
        return None

    def _plan_transpose_fusion(self, fetches):
        """The NCHW flatten / un-flatten around the dense layers that follow or precede an image stack
        (`tf.reshape(output, [-1, 4*4*4*DIM])` then `tf.concat([output, z_output], 1)`, gmgan_inference_cifar10.py:292-294) is a
        batched 2-D transpose with a copy on one side: transpose -> copy into the concat's columns (forward); column slice ->
        transpose -> activation gradient (backward: three dependent launches on the step's critical chain).  gg_transpose_b2d_ex
        reads batches at a row stride (the slice disappears), writes them at a row stride (the copy disappears) and can
        multiply by act'(y) on the way out (the activation-gradient launch disappears)."""
        self.tr_fuse, self.virtual_slice, self.col_placed, self.gather_add = {}, set(), {}, {}
        if os.environ.get("GG_FUSE_TRANSPOSE", "1") == "0":
            return
        uses = self._use_counts(fetches)
        pos = {n.id: i for i, n in enumerate(self.order)}
        consumers = {}
        for n in self.order:
            if n.id in self.fed:
                continue
            for i in n.inputs:
                consumers.setdefault(i.id, []).append(n)
        placed = set(self.placed) | set(self.placed_flat)
        # one_hot(idx) @ table + x  ->  the row gather adds x on its way out (the add launch disappears from the head of the chain)
        if os.environ.get("GG_GATHER", "1") != "0":
            for m in self.order:
                if m.op != "matmul" or m.id in self.fed or m.inputs[0].op != "one_hot" or m.inputs[0].id in self.fed or len(m.inputs) != 2 \
                        or m.attrs["ta"] or m.attrs["tb"] or m.attrs["act"] is not None or uses.get(m.id, 0) != 1 or m.id in placed \
                        or m.id in self.fuse_mask or len(consumers.get(m.id, ())) != 1:
                    continue
                c = consumers[m.id][0]
                if c.op != "binary" or c.attrs["fn"] != "add" or c.id in self.fed or c.id in self.fused_alias or c.inputs[0] is not m:
                    continue
                other = c.inputs[1]
                if tuple(other.shape) != tuple(m.shape) or tuple(c.shape) != tuple(m.shape) or other.dtype != float32 \
                        or other.id not in pos or self._is_ancestor(m, other):
                    continue
                self.gather_add[m.id] = (c, other)
                self.fused_alias[c.id] = m.id
        for t in self.order:
            if t.op != "transpose" or t.id in self.fed or t.dtype != float32 or t.id in placed:
                continue
            form = self._b2d_form(t)
            if form is None:
                continue
            Bt, R, Cc = form
            if Bt > 65535:
                continue
            info = dict(x_node=None, x_off=0, x_bs=R * Cc, y_concat=None, y_off=0, y_bs=R * Cc, mask=None)
            # input side: a column slice of a [Bt, total] matrix behind single-use reshapes
            x, chain = t.inputs[0], []
            while x.op == "reshape" and x.id not in self.fed and uses.get(x.id, 0) == 1:
                chain.append(x)
                x = x.inputs[0]
            if x.op == "slice" and x.id not in self.fed and uses.get(x.id, 0) == 1 and x.id not in placed and not self._slice_is_view(x):
                src = x.inputs[0]
                if len(src.shape) == 2 and x.attrs["axis"] == 1 and src.shape[0] == Bt and x.attrs["size"] == R * Cc and src.dtype == float32:
                    info.update(x_node=src, x_off=x.attrs["start"], x_bs=src.shape[1], slice=x)
            # output side: the only reader is an activation gradient, or (through single-use reshapes) a column concat
            if uses.get(t.id, 0) == 1 and len(consumers.get(t.id, ())) == 1:
                c = consumers[t.id][0]
                if c.op == "binary" and c.attrs.get("fn") in ("leaky_grad", "relu_grad", "tanh_grad", "sigmoid_grad") and c.inputs[1] is t \
                        and c.id not in self.fed and c.id not in placed and c.id not in self.fused_alias \
                        and tuple(c.inputs[0].shape) == tuple(t.shape) and tuple(c.shape) == tuple(t.shape) \
                        and c.inputs[0].id in pos and not self._is_ancestor(t, c.inputs[0]):
                    info["mask"] = (c.inputs[0], c.attrs["fn"][:-5], float(c.attrs["alpha"]), c)
                else:
                    while c.op == "reshape" and c.id not in self.fed and uses.get(c.id, 0) == 1 and len(consumers.get(c.id, ())) == 1:
                        c = consumers[c.id][0]
                    if c.op == "concat" and c.id not in self.fed and c.dtype == float32 and len(c.shape) == 2 and c.attrs["axis"] == 1 \
                            and c.shape[0] == Bt and c.id not in self.inplace_concat and c.id not in placed:
                        off = 0
                        for inp in c.inputs:
                            own = inp
                            while own.op == "reshape" and own is not t:
                                own = own.inputs[0]
                            if own is t and inp.size == Bt * R * Cc:
                                info.update(y_concat=c, y_off=off, y_bs=c.shape[1])
                                break
                            off += inp.shape[1]
            if info["x_node"] is None and info["y_concat"] is None and info["mask"] is None:
                continue
            self.tr_fuse[t.id] = info
            if info["x_node"] is not None:
                self.virtual_slice.add(info["slice"].id)
            if info["mask"] is not None:
                y, act, alpha, c = info["mask"]
                self.fuse_mask[t.id] = (y, act, alpha)
                self.fused_alias[c.id] = t.id
            if info["y_concat"] is not None:
                self.col_placed.setdefault(info["y_concat"].id, {})[info["y_off"]] = t.id

    # ---- element-wise cluster fusion (gg/fuse.py) ------------------------------------------------------------------------
    def _plan_ew_fusion(self, fetches):
        """group connected element-wise nodes into clusters that run as one gg_ew_run launch each, and re-sort the plan so
        that every cluster sits behind all of its inputs (GG_FUSE_EW=0: one launch per node, as before)"""
        from . import fuse
        self.ew_cluster_of, self.ew_clusters = {}, []
        if os.environ.get("GG_FUSE_EW", "1") == "0":
            return
        ext = {}
        for f in fetches:
            if isinstance(f, Tensor):
                ext[f.id] = ext.get(f.id, 0) + 1
            else:
                for d in self._op_roots(f):
                    ext[d.id] = ext.get(d.id, 0) + 1
        for gid, (y, _act, _alpha) in self.fuse_mask.items():
            ext[y.id] = ext.get(y.id, 0) + 1
        for nid, (_c, other) in self.gather_add.items():
            ext[other.id] = ext.get(other.id, 0) + 1
        excluded = set(self.fused_alias)
        node_of = {n.id: n for n in self.order}
        edges = [(y, node_of[gid]) for gid, (y, _a, _al) in self.fuse_mask.items() if gid in node_of]
        edges += [(other, node_of[mid]) for mid, (_c, other) in self.gather_add.items() if mid in node_of]
        clusters = fuse.Planner(self.order, self.fed, ext, excluded, edges).build()
        for cl in clusters:
            for n in cl.nodes():
                assert n.id not in self.ew_cluster_of, "node %s in two element-wise clusters" % n
                self.ew_cluster_of[n.id] = cl
        self.ew_clusters = clusters

    def _resort(self):
        """stable topological re-sort of the plan with every element-wise cluster as ONE unit (keyed by its last node's old
        position) and with the operands the peepholes added (activation masks, the addend of a gather) as dependencies"""
        pos = {n.id: i for i, n in enumerate(self.order)}
        unit_of, unit_nodes, unit_key = {}, {}, {}
        for n in self.order:
            cl = self.ew_cluster_of.get(n.id)
            u = ("c", id(cl)) if cl is not None else ("n", n.id)
            unit_of[n.id] = u
            unit_nodes.setdefault(u, []).append(n)
            unit_key[u] = max(unit_key.get(u, -1), pos[n.id])
        extra = {}
        for gid, (y, _act, _alpha) in self.fuse_mask.items():     # the fused dgrad / transpose launch reads its activation mask y
            extra.setdefault(gid, []).append(y)
        for mid, (_c, other) in self.gather_add.items():
            extra.setdefault(mid, []).append(other)
        if not self.ew_clusters and not extra:
            return
        deps, succ = {u: set() for u in unit_nodes}, {u: set() for u in unit_nodes}
        for n in self.order:
            if n.id in self.fed:
                continue
            for i in list(n.inputs) + extra.get(n.id, []):
                if i.id in unit_of and unit_of[i.id] != unit_of[n.id]:
                    deps[unit_of[n.id]].add(unit_of[i.id])
                    succ[unit_of[i.id]].add(unit_of[n.id])
        import heapq
        ready = [(unit_key[u], u) for u in unit_nodes if not deps[u]]
        heapq.heapify(ready)
        left = {u: len(deps[u]) for u in unit_nodes}
        order = []
        while ready:
            _, u = heapq.heappop(ready)
            order.extend(unit_nodes[u])
            for v in succ[u]:
                left[v] -= 1
                if left[v] == 0:
                    heapq.heappush(ready, (unit_key[v], v))
        assert len(order) == len(self.order), "element-wise clustering produced a dependency cycle"
        self.order = order

    def _emit_ew_cluster(self, cl):
        from . import fuse
        d = cl.desc
        s0 = len(self.steps)
        out_bufs = []
        for m in d["out_nodes"]:
            target = cl.reduce if (d["reduce"] and m is cl.members[-1]) else m
            out_bufs.append(self._alloc(target))
        for m in d["interior"]:
            if m.id not in self.buf:
                self.buf[m.id] = None               # lives in a register of the fused launch only
        for a in cl.aliases:                        # in plan order: an alias of an alias resolves through its input
            self.buf[a.id] = self.buf[a.inputs[0].id]
        in_ptrs = [self.buf[ld["node"].id].data_ptr() for ld in d["loads"]]
        prog = fuse.to_struct(d, in_ptrs, [t.data_ptr() for t in out_bufs])
        self.keep.append(prog)
        ref = C.byref(prog)
        self.steps.append(lambda st: cabi.call("gg_ew_run", ref, st))
        cl.emitted = True
        reads = set()
        for ld in d["loads"]:
            reads |= self._owners(ld["node"])
        for n in cl.nodes():
            self.owner[n.id] = frozenset([cl.key])
        self._add_groups(s0, len(self.steps), reads, cl.key, barrier=False, node=(cl.reduce if cl.reduce is not None else cl.members[-1]))

    def _late_vars(self, pairs, force=False):
        """ids of the variables whose gradients become ready last (within one tensor-core launch of the deepest one): the
        optimiser update is split so that only THEIR update waits for the end of the backward pass; every other parameter is
        updated as soon as its own gradient and the last kernel reading it are done.  OPT-IN (GG_SPLIT_UPDATE=1): measured on the
        B200 the early update (87-114 MB of L2 traffic) slows the last backward kernels it overlaps by more than it saves — 0.888
        vs 0.874 ms per iteration (profiles/sched_variants_r2.txt)."""
        if (os.environ.get("GG_SPLIT_UPDATE", "0") != "1" and not force) or len(pairs) < 2:
            return set()
        pos = getattr(self, "ready_pos", {})
        ready = {}
        for v, g in pairs:
            own = g
            while own.op in ("reshape", "stop_gradient") and own.id not in self.fed and own.inputs:
                own = own.inputs[0]
            ready[v.id] = pos.get(own.id, 0)
        top = max(ready.values())
        late = set(vid for vid, r in ready.items() if r > top - 6)
        if len(late) == len(pairs):
            return set()
        return late

    def _concat_storage(self, cid):
        if cid not in self._concat_bufs:
            node = next(n for n in self.order if n.id == cid)
            self._concat_bufs[cid] = self.rt.empty(node.shape, node.dtype)
        return self._concat_bufs[cid]

    def _ws(self, nbytes):
        torch = _torch()
        # zero-initialised once: the tensor-core conv kernels keep self-cleaning arrival tickets at its start
        return torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=self.rt.dev())

    def _emit(self, node):
        if node.id in self.fed:
            self.buf[node.id] = self.rt.feed_buffer(node)[0]
            return
        fn = getattr(self, "_emit_" + node.op, None)
        if fn is None:
            raise NotImplementedError("no launcher for op %r" % node.op)
        fn(node)

    def _in(self, node, i):
        return self.buf[node.inputs[i].id]

    # leaves
    def _emit_placeholder(self, node):
        raise cabi.GGError("placeholder %s must be fed" % node.name)

    def _emit_const(self, node):
        self.buf[node.id] = self.rt.const_buffer(node)

    def _emit_param(self, node):
        self.buf[node.id] = self.rt.param_buffer(node)

    # aliases
    def _emit_reshape(self, node):
        self.buf[node.id] = self._in(node, 0)

    _emit_stop_gradient = _emit_reshape

    def _emit_aux(self, node):
        self.buf[node.id] = self.extra[(node.inputs[0].id, node.attrs["k"])]

    # element-wise
    def _emit_unary(self, node):
        x, y = self._in(node, 0), self._alloc(node)
        op, a, b, n = cabi.UNARY[node.attrs["fn"]], node.attrs["a"], node.attrs["b"], node.size
        xp, yp = x.data_ptr(), y.data_ptr()
        self.steps.append(lambda st: cabi.call("gg_unary", op, xp, yp, n, a, b, st))

    @staticmethod
    def _bcast_strides(shape, out_shape):
        nd = len(out_shape)
        shape = (1,) * (nd - len(shape)) + tuple(shape)
        strides, acc = [0] * nd, 1
        for i in range(nd - 1, -1, -1):
            strides[i] = 0 if shape[i] == 1 and out_shape[i] != 1 else acc
            acc *= shape[i]
        for i in range(nd):
            if shape[i] == 1:
                strides[i] = 0
        return strides

    @staticmethod
    def _merge_dims(dims, sa, sb):
        """merge adjacent dims that are contiguous in both operands; drop size-1 dims; pad to 4"""
        d, a, b = [], [], []
        for n, x, y in zip(dims, sa, sb):
            if n == 1:
                continue
            if d and a[-1] == x * n and b[-1] == y * n:
                d[-1] *= n
                a[-1], b[-1] = x, y
            else:
                d.append(n); a.append(x); b.append(y)
        if len(d) > 4:
            raise NotImplementedError("broadcast pattern needs more than 4 dims: %s" % (dims,))
        while len(d) < 4:
            d.insert(0, 1); a.insert(0, 0); b.insert(0, 0)
        return d, a, b

    def _launch_binary(self, fn, alpha, ap, a_shape, bp, b_shape, op_, out_shape):
        sa = self._bcast_strides(a_shape, out_shape)
        sb = self._bcast_strides(b_shape, out_shape)
        d, a4, b4 = self._merge_dims(list(out_shape), sa, sb)
        dims, sa4, sb4 = cabi.int4(d), cabi.int4(a4), cabi.int4(b4)
        self.keep.append((dims, sa4, sb4))
        code = cabi.BINARY[fn]
        self.steps.append(lambda st: cabi.call("gg_binary", code, ap, bp, op_, dims, sa4, sb4, alpha, st))

    def _emit_binary(self, node):
        if node.id in self.fused_alias:                   # computed by the producing dgrad launch (_plan_actgrad_fusion)
            self.buf[node.id] = self.buf[self.fused_alias[node.id]]
            return
        a, b, out = self._in(node, 0), self._in(node, 1), self._alloc(node)
        self._launch_binary(node.attrs["fn"], node.attrs["alpha"], a.data_ptr(), node.inputs[0].shape, b.data_ptr(),
                            node.inputs[1].shape, out.data_ptr(), tuple(node.shape))

    def _emit_broadcast(self, node):
        x, out = self._in(node, 0), self._alloc(node)
        self._launch_binary("add", 0.0, x.data_ptr(), node.inputs[0].shape, self.rt.zero().data_ptr(), (1,), out.data_ptr(),
                            tuple(node.shape))

    def _emit_add_n(self, node):
        out = self._alloc(node)
        ptrs = (C.c_void_p * len(node.inputs))(*[self._in(node, i).data_ptr() for i in range(len(node.inputs))])
        self.keep.append(ptrs)
        cnt, n, op_ = len(node.inputs), node.size, out.data_ptr()
        self.steps.append(lambda st: cabi.call("gg_add_n", ptrs, cnt, op_, n, st))

    def _emit_cast(self, node):
        x, y = self._in(node, 0), self._alloc(node)
        xp, yp, n = x.data_ptr(), y.data_ptr(), node.size
        if node.dtype == float32:
            self.steps.append(lambda st: cabi.call("gg_cast_i32_f32", xp, yp, n, 1.0, 0.0, st))
        else:
            self.steps.append(lambda st: cabi.call("gg_cast_f32_i32", xp, yp, n, st))

    def _emit_reduce(self, node):
        x, y = self._in(node, 0), self._alloc(node)
        shp = node.inputs[0].shape
        axes = node.attrs["axes"]
        outer, red, inner = prod(shp[:axes[0]]), prod(shp[axes[0]:axes[-1] + 1]), prod(shp[axes[-1] + 1:])
        code, xp, yp = cabi.REDUCE[node.attrs["fn"]], x.data_ptr(), y.data_ptr()
        need = cabi.lib.gg_reduce_workspace(outer, red, inner)
        if need and red >= 256:
            ws = self._ws(need)
            self.keep.append(ws)
            wp, wn = ws.data_ptr(), ws.numel()
            self.steps.append(lambda st: cabi.call("gg_reduce_ws", code, xp, yp, outer, red, inner, wp, wn, st))
        else:
            self.steps.append(lambda st: cabi.call("gg_reduce", code, xp, yp, outer, red, inner, st))

    def _emit_softmax(self, node):
        x, y = self._in(node, 0), self._alloc(node)
        R, Cc, xp, yp = prod(node.shape[:-1]), node.shape[-1], x.data_ptr(), y.data_ptr()
        self.steps.append(lambda st: cabi.call("gg_softmax_fwd", xp, yp, R, Cc, st))

    def _emit_softmax_grad(self, node):
        y, g, out = self._in(node, 0), self._in(node, 1), self._alloc(node)
        R, Cc, yp, gp, op_ = prod(node.shape[:-1]), node.shape[-1], y.data_ptr(), g.data_ptr(), out.data_ptr()
        self.steps.append(lambda st: cabi.call("gg_softmax_bwd", yp, gp, op_, R, Cc, st))

    def _emit_argmax(self, node):
        x, y = self._in(node, 0), self._alloc(node)
        shp = node.inputs[0].shape
        R, Cc, xp, yp = prod(shp[:-1]), shp[-1], x.data_ptr(), y.data_ptr()
        self.steps.append(lambda st: cabi.call("gg_argmax", xp, yp, R, Cc, st))

    def _emit_one_hot(self, node):
        x, y = self._in(node, 0), self._alloc(node)
        n, depth, xp, yp = node.inputs[0].size, node.attrs["depth"], x.data_ptr(), y.data_ptr()
        self.steps.append(lambda st: cabi.call("gg_one_hot", xp, yp, n, depth, st))

    def _emit_random(self, node):
        out = self._alloc(node)
        self.has_random = True
        seed, sid, tick, op_, n = self.rt.seed, node.id & 0xFFFFFFFF, self.rt.tick().data_ptr(), out.data_ptr(), node.size
        kind, a, b = node.attrs["kind"], node.attrs.get("a", 0.0), node.attrs.get("b", 1.0)
        rt = self.rt
        if kind == "normal":
            self.steps.append(lambda st: cabi.call("gg_rng_normal", op_, n, a, b, rt.rng_seed(), sid, tick, st))
        elif kind == "uniform":
            self.steps.append(lambda st: cabi.call("gg_rng_uniform", op_, n, a, b, rt.rng_seed(), sid, tick, st))
        else:
            pp, K = self._in(node, 0).data_ptr(), node.inputs[0].size
            self.steps.append(lambda st: cabi.call("gg_rng_categorical", op_, n, pp, K, rt.rng_seed(), sid, tick, st))

    # layout
    def _emit_transpose(self, node):
        info = self.tr_fuse.get(node.id)
        if info is not None:
            Bt, R, Cc = self._b2d_form(node)
            if info["x_node"] is not None:
                xp = self.buf[info["x_node"].id].data_ptr() + 4 * info["x_off"]
            else:
                xp = self._in(node, 0).data_ptr()
            if info["y_concat"] is not None:
                store = self._concat_storage(info["y_concat"].id)
                self.buf[node.id] = store[info["y_off"]:]            # rows of the concat at a column offset: only the concat reads it
                yp = store.data_ptr() + 4 * info["y_off"]
            else:
                yp = self._alloc(node).data_ptr()
            mp, mcode, malpha = None, 0, 0.0
            if info["mask"] is not None:
                ynode, act, malpha, _c = info["mask"]
                mp, mcode = self.buf[ynode.id].data_ptr(), cabi.ACT[act]
            x_bs, y_bs = info["x_bs"], info["y_bs"]
            self.steps.append(lambda st: cabi.call("gg_transpose_b2d_ex", xp, yp, Bt, R, Cc, x_bs, y_bs, mp, mcode, malpha, st))
            return
        x, y = self._in(node, 0), self._alloc(node)
        shp, perm = list(node.inputs[0].shape), list(node.attrs["perm"])
        xp, yp = x.data_ptr(), y.data_ptr()
        nd = len(shp)
        # batched 2-D transpose fast path: perm = (0, 2..n-1, 1) or (0, n-1, 1..n-2)
        if nd >= 3 and perm[0] == 0 and perm[1:] == list(range(2, nd)) + [1]:
            Bt, R, Cc = shp[0], shp[1], prod(shp[2:])          # [B, R, C] -> [B, C, R]
            self.steps.append(lambda st: cabi.call("gg_transpose_b2d", xp, yp, Bt, R, Cc, st))
            return
        if nd >= 3 and perm[0] == 0 and perm[1:] == [nd - 1] + list(range(1, nd - 1)):
            Bt, R, Cc = shp[0], prod(shp[1:nd - 1]), shp[nd - 1]
            self.steps.append(lambda st: cabi.call("gg_transpose_b2d", xp, yp, Bt, R, Cc, st))
            return
        if nd == 2:
            R, Cc = shp
            self.steps.append(lambda st: cabi.call("gg_transpose_b2d", xp, yp, 1, R, Cc, st))
            return
        if nd > 4:
            raise NotImplementedError("transpose of rank %d" % nd)
        pad = 4 - nd
        dims = cabi.int4([1] * pad + shp)
        p4 = cabi.int4(list(range(pad)) + [p + pad for p in perm])
        self.keep.append((dims, p4))
        self.steps.append(lambda st: cabi.call("gg_transpose4", xp, yp, dims, p4, st))

    def _emit_concat(self, node):
        if node.id in self.inplace_concat:
            self.buf[node.id] = self._concat_storage(node.id)      # the pieces were written in place by their producers
            return
        pre = self.col_placed.get(node.id, {})       # pieces a strided transpose already wrote in place (_plan_transpose_fusion)
        if pre:
            out = self.buf[node.id] = self._concat_storage(node.id)
        else:
            out = self._alloc(node)
        axis = node.attrs["axis"]
        outer = prod(node.shape[:axis])
        inner = prod(node.shape[axis + 1:])
        dst_ld = node.shape[axis] * inner
        off = 0
        if node.dtype != float32:
            raise NotImplementedError("concat of int tensors")
        # one scheduling group per piece: a copy waits for ITS piece only, and a reader of the concat waits for all of them —
        # the copy of an early piece (the z branch of `tf.concat([output, z_output], 1)`) leaves the critical chain
        keys = set()
        for i, inp in enumerate(node.inputs):
            cols = inp.shape[axis] * inner
            if off in pre:
                keys |= self._owners(inp)
            else:
                sp, dp = self._in(node, i).data_ptr(), out.data_ptr() + off * 4
                s0 = len(self.steps)
                self.steps.append(lambda st, sp=sp, dp=dp, cols=cols: cabi.call("gg_copy2d", sp, cols, dp, dst_ld, outer, cols, 0, st))
                key = "%d#piece%d" % (node.id, i)
                self._add_groups(s0, len(self.steps), set(self._owners(inp)), key, barrier=False, node=node)
                keys.add(key)
            off += cols
        self.owner[node.id] = frozenset(keys)
        self.self_grouped.add(node.id)

    def _emit_slice(self, node):
        axis, start, size = node.attrs["axis"], node.attrs["start"], node.attrs["size"]
        shp = node.inputs[0].shape
        if self._slice_is_view(node) and node.id not in self.placed and os.environ.get("GG_INPLACE_CONCAT", "1") != "0":
            inner = prod(shp[axis + 1:])
            self.buf[node.id] = self._in(node, 0)[start * inner:(start + size) * inner]   # rows of a contiguous tensor: a view
            return
        if node.id in self.virtual_slice:           # read in place by the strided transpose that consumes it
            self.buf[node.id] = None
            return
        x, out = self._in(node, 0), self._alloc(node)
        outer, inner = prod(shp[:axis]), prod(shp[axis + 1:])
        src_ld, cols = shp[axis] * inner, size * inner
        sp, dp = x.data_ptr() + start * inner * 4, out.data_ptr()
        self.steps.append(lambda st: cabi.call("gg_copy2d", sp, src_ld, dp, cols, outer, cols, 0, st))

    def _emit_pad(self, node):
        x, out = self._in(node, 0), self._alloc(node)
        axis, start, total = node.attrs["axis"], node.attrs["start"], node.attrs["total"]
        shp = node.inputs[0].shape
        outer, inner = prod(shp[:axis]), prod(shp[axis + 1:])
        cols, dst_ld = shp[axis] * inner, total * inner
        sp, dp0, dp, n = x.data_ptr(), out.data_ptr(), out.data_ptr() + start * inner * 4, node.size
        self.steps.append(lambda st: cabi.call("gg_fill", dp0, n, 0.0, st))
        self.steps.append(lambda st: cabi.call("gg_copy2d", sp, cols, dp, dst_ld, outer, cols, 0, st))

    def _emit_tile(self, node):
        x, out = self._in(node, 0), self._alloc(node)
        shp, mult = list(node.inputs[0].shape), list(node.attrs["multiples"])
        # out viewed as [m0, s0, m1, s1, ...]; x broadcast over the m axes
        out_view, x_view = [], []
        for s, m in zip(shp, mult):
            out_view += [m, s]
            x_view += [1, s]
        self._launch_binary("add", 0.0, x.data_ptr(), tuple(x_view), self.rt.zero().data_ptr(), (1,), out.data_ptr(),
                            tuple(out_view))

    # dense / conv / bn
    def _emit_matmul(self, node):
        oh = node.inputs[0]
        if oh.op == "one_hot" and oh.id not in self.fed and len(node.inputs) == 2 and not node.attrs["ta"] and not node.attrs["tb"] \
                and node.attrs["act"] is None and node.id not in self.fuse_mask and os.environ.get("GG_GATHER", "1") != "0" \
                and len(oh.shape) == 2:
            # one_hot(idx) @ table is a row gather: no one-hot matrix, no GEMM, and the launch waits for idx, not for one_hot
            idx, table = self.buf[oh.inputs[0].id], self._in(node, 1)
            M, N = node.shape
            extra = self.gather_add.get(node.id)           # (the add node, its other operand): mu_k + noise in the same launch
            out = self._alloc(extra[0] if extra else node)
            self.buf[node.id] = out
            adp = self.buf[extra[1].id].data_ptr() if extra else None
            ip, tp, op_, depth = idx.data_ptr(), table.data_ptr(), out.data_ptr(), oh.attrs["depth"]
            self.steps.append(lambda st: cabi.call("gg_gather_rows", ip, tp, adp, op_, M, N, depth, st))
            self.read_override[node.id] = [oh.inputs[0], node.inputs[1]] + ([extra[1]] if extra else [])
            return
        a, b = self._in(node, 0), self._in(node, 1)
        bias = self._in(node, 2).data_ptr() if len(node.inputs) == 3 else None
        out = self._alloc(node)
        ta, tb = int(node.attrs["ta"]), int(node.attrs["tb"])
        M, N = node.shape
        K = node.inputs[0].shape[0] if ta else node.inputs[0].shape[1]
        ws = self._ws(cabi.lib.gg_gemm_workspace(M, N, K))
        self.keep.append(ws)
        act, alpha = cabi.ACT[node.attrs["act"]], node.attrs["alpha"]
        ap, bp, op_, wp, wn = a.data_ptr(), b.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel()
        if node.id in self.bn_stats:
            self._emit_bnstats_launch(node, ap, bp, bias, op_, act, alpha, wp, wn)
            return
        if node.id in self.fuse_mask:                     # dx = act'(y) * (dy W^T): the dense dgrad + activation gradient, one launch
            y, mact, malpha = self.fuse_mask[node.id]
            yp, mcode = self.buf[y.id].data_ptr(), cabi.ACT[mact]
            self.steps.append(lambda st: cabi.call("gg_conv2d_dgrad_actgrad", ap, bp, op_, yp, mcode, malpha, M, 1, 1, N, K, 1, 1, 0, 0,
                                                   1, 1, wp, wn, st))
            return
        self.steps.append(lambda st: cabi.call("gg_gemm", ap, bp, bias, op_, M, N, K, ta, tb, act, alpha, wp, wn, st))

    def _emit_conv(self, node):
        g = node.attrs
        geo = (g["B"], g["H"], g["W"], g["Ci"], g["Co"], g["k"], g["stride"], g["pad_t"], g["pad_l"], g["Ho"], g["Wo"])
        a, b = self._in(node, 0), self._in(node, 1)
        out = self._alloc(node)
        ap, bp, op_ = a.data_ptr(), b.data_ptr(), out.data_ptr()
        mode = g["mode"]
        if mode == "wgrad":
            need = cabi.lib.gg_conv2d_wgrad_workspace(g["B"], g["H"], g["W"], g["Ci"], g["Co"], g["k"], g["stride"], g["Ho"], g["Wo"])
            ws = self._ws(need)
            self.keep.append(ws)
            wp, wn = ws.data_ptr(), ws.numel()
            self.steps.append(lambda st: cabi.call("gg_conv2d_wgrad", ap, bp, op_, *geo, wp, wn, st))
            return
        bias = self._in(node, 2).data_ptr() if len(node.inputs) == 3 else None
        act, alpha = cabi.ACT[g["act"]], g["alpha"]
        ws = self._ws(cabi.lib.gg_conv2d_workspace(0 if mode == "fwd" else 1, g["B"], g["H"], g["W"], g["Ci"], g["Co"], g["k"],
                                                   g["stride"], g["Ho"], g["Wo"]))
        self.keep.append(ws)
        wp, wn = ws.data_ptr(), ws.numel()
        if node.id in self.fuse_mask:                     # dgrad + the activation gradient that follows it, one launch
            y, mact, malpha = self.fuse_mask[node.id]
            yp, mcode = self.buf[y.id].data_ptr(), cabi.ACT[mact]
            self.steps.append(lambda st: cabi.call("gg_conv2d_dgrad_actgrad", ap, bp, op_, yp, mcode, malpha, *geo, wp, wn, st))
            return
        if node.id in self.bn_stats:                      # + the statistics of the batch norm that follows (_plan_bn_stats)
            self._emit_bnstats_launch(node, ap, bp, bias, op_, act, alpha, wp, wn)
            return
        name = "gg_conv2d_fwd" if mode == "fwd" else "gg_conv2d_dgrad"
        self.steps.append(lambda st: cabi.call(name, ap, bp, bias, op_, *geo, act, alpha, wp, wn, st))

    def _emit_bnstats_launch(self, node, ap, bp, bias, op_, act, alpha, wp, wn):
        info = self.bn_stats[node.id]
        info["buf"] = self.rt.empty((info["tiles"], 2, info["C"]))
        sp, mode, geo = info["buf"].data_ptr(), info["mode"], info["geo"]
        self.steps.append(lambda st: cabi.call("gg_conv2d_bnstats", mode, ap, bp, bias, op_, sp, *geo, act, alpha, wp, wn, st))

    def _emit_bn(self, node):
        torch = _torch()
        x = self._in(node, 0)
        gamma, beta = self._in(node, 1), self._in(node, 2)
        y = self._alloc(node)
        Cc = node.shape[-1]
        R = node.size // Cc
        mean, rstd = self.rt.empty((Cc,)), self.rt.empty((Cc,))
        self.extra[(node.id, 1)], self.extra[(node.id, 2)] = mean, rstd
        act, alpha, eps = cabi.ACT[node.attrs["act"]], node.attrs["alpha"], node.attrs["eps"]
        world = ggdist.world_size()
        if node.id in self.bn_from_stats:
            # the producing conv / dense launch left per-m-tile sums: fold + normalise + activation in one element-wise pass
            info = self.bn_stats[self.bn_from_stats[node.id]]
            xp, pp, gp, bp, yp, mp, rp = (t.data_ptr() for t in (x, info["buf"], gamma, beta, y, mean, rstd))
            S = info["tiles"]
            self.steps.append(lambda st: cabi.call("gg_bn_apply", xp, pp, S, float(R), gp, bp, eps, yp, mp, rp, R, Cc, act, alpha, st))
            return
        if self._bn_fused(R, Cc, world):
            # statistics + normalise + activation in ONE launch (no cross-rank exchange needed)
            xp, gp, bp, yp, mp, rp = (t.data_ptr() for t in (x, gamma, beta, y, mean, rstd))
            self.steps.append(lambda st: cabi.call("gg_bn_fwd_fused", xp, gp, bp, eps, yp, mp, rp, R, Cc, act, alpha, st))
            return
        arena = self._bn_dp_arena(R, Cc, world)
        if arena is not None:
            # SyncBN in ONE launch: the per-channel sums are totalled over the ranks inside the kernel (NVLink peer stores +
            # epoch flags in the exchange arena); one region per call site
            site = arena.alloc(cabi.lib.gg_bn_dp_site_bytes(Cc, world))
            xp, gp, bp, yp, mp, rp = (t.data_ptr() for t in (x, gamma, beta, y, mean, rstd))
            peers, rk = arena.peers, ggdist.rank()
            fn = lambda st: cabi.call("gg_bn_fwd_fused_dp", xp, gp, bp, eps, yp, mp, rp, R, Cc, act, alpha, peers, rk, world, site, st)
            fn.is_small_allreduce = self._bn_dp_ordered(R, Cc)
            self.steps.append(fn)
            return
        S = cabi.lib.gg_bn_slices(R, Cc)
        part = self.rt.empty((S, 2, Cc))
        self.keep.append(part)
        xp, pp, gp, bp, yp, mp, rp = (t.data_ptr() for t in (x, part, gamma, beta, y, mean, rstd))
        self.steps.append(lambda st: cabi.call("gg_bn_stats", xp, pp, R, Cc, st))
        if world > 1 and self.rt.sync_bn:
            folded = self.rt.empty((2, Cc))
            self.keep.append(folded)
            fp = folded.data_ptr()
            self.steps.append(lambda st: cabi.call("gg_bn_fold_partials", pp, S, fp, Cc, st))
            self._all_reduce_small(folded, 2 * Cc)
            cnt = float(R * world)
            self.steps.append(lambda st: cabi.call("gg_bn_apply", xp, fp, 1, cnt, gp, bp, eps, yp, mp, rp, R, Cc, act, alpha, st))
        else:
            self.steps.append(lambda st: cabi.call("gg_bn_apply", xp, pp, S, float(R), gp, bp, eps, yp, mp, rp, R, Cc, act, alpha, st))

    def _bn_fused(self, R, Cc, world):
        return (world == 1 or not self.rt.sync_bn) and os.environ.get("GG_BN_FUSED", "1") != "0" and \
            cabi.lib.gg_bn_fused_supported(R, Cc) == 1

    @staticmethod
    def _bn_dp_ordered(R, Cc):
        """Must the one-launch SyncBN kernels run one at a time in list order (scheduler chain "peer")?  Every call site has
        its own flags, so concurrent kernels cannot confuse each other; the only hazard is residency: rank A spinning in
        kernel X while rank B's SMs are full of kernel Y's spinning CTAs.  A kernel of <= 148 CTAs (512 threads, 2+ CTAs per
        SM) can never fill a GPU, and a training graph has at most two independent batch-norm chains (generator /
        extractor), so small grids run unordered — chaining them serialised the two chains and cost ~0.2 ms per iteration
        at N=2 (profiles/dp2_variants_r2.txt).  GG_BN_DP_ORDER=1 forces the chain."""
        if os.environ.get("GG_BN_DP_ORDER", "0") == "1":
            return True
        return cabi.lib.gg_bn_fused_grid(R, Cc) > 148

    def _bn_dp_arena(self, R, Cc, world):
        """the peer exchange arena when this batch norm can run as the one-launch data-parallel kernel, else None"""
        if world <= 1 or not self.rt.sync_bn or os.environ.get("GG_BN_FUSED", "1") == "0" or \
                cabi.lib.gg_bn_fused_supported(R, Cc) != 1:
            return None
        return ggdist.peer_arena()

    def _emit_bn_grad(self, node):
        gy, x, y, mean, rstd, gamma = (self._in(node, i) for i in range(6))
        dx = self._alloc(node)
        Cc = node.shape[-1]
        R = node.size // Cc
        dgb = self.rt.empty((2, Cc))          # [dbeta ; dgamma]
        self.extra[(node.id, 1)] = dgb[Cc:]   # dgamma
        self.extra[(node.id, 2)] = dgb[:Cc]   # dbeta
        act, alpha = cabi.ACT[node.attrs["act"]], node.attrs["alpha"]
        if self._bn_fused(R, Cc, ggdist.world_size()):
            gyp, xp, yp, mp, rp, gp, dxp = (t.data_ptr() for t in (gy, x, y, mean, rstd, gamma, dx))
            dgp, dbp = dgb.data_ptr() + 4 * Cc, dgb.data_ptr()
            self.steps.append(lambda st: cabi.call("gg_bn_bwd_fused", gyp, xp, yp, mp, rp, gp, dxp, dgp, dbp, R, Cc, act, alpha, st))
            return
        world = ggdist.world_size()
        arena = self._bn_dp_arena(R, Cc, world)
        if arena is not None:
            site = arena.alloc(cabi.lib.gg_bn_dp_site_bytes(Cc, world))
            gyp, xp, yp, mp, rp, gp, dxp = (t.data_ptr() for t in (gy, x, y, mean, rstd, gamma, dx))
            dgp, dbp = dgb.data_ptr() + 4 * Cc, dgb.data_ptr()
            peers, rk = arena.peers, ggdist.rank()
            fn = lambda st: cabi.call("gg_bn_bwd_fused_dp", gyp, xp, yp, mp, rp, gp, dxp, dgp, dbp, R, Cc, act, alpha, peers, rk,
                                      world, site, st)
            fn.is_small_allreduce = self._bn_dp_ordered(R, Cc)
            self.steps.append(fn)
            return
        S = cabi.lib.gg_bn_slices(R, Cc)
        part = self.rt.empty((S, 2, Cc))
        self.keep.append(part)
        gyp, xp, yp, mp, rp, gp, pp, dxp, dgbp = (t.data_ptr() for t in (gy, x, y, mean, rstd, gamma, part, dx, dgb))
        self.steps.append(lambda st: cabi.call("gg_bn_bwd_reduce", gyp, xp, yp, mp, rp, gp, None, pp, R, Cc, act, alpha, st))
        # local sums are the (local) parameter gradients; the data-parallel all-reduce of the gradient bucket sums them
        self.steps.append(lambda st: cabi.call("gg_bn_fold_partials", pp, S, dgbp, Cc, st))
        world = ggdist.world_size()
        if world > 1 and self.rt.sync_bn:
            glob = self.rt.empty((2, Cc))
            self.keep.append(glob)
            glp = glob.data_ptr()
            self.steps.append(lambda st: cabi.call("gg_unary", cabi.UNARY["copy"], dgbp, glp, 2 * Cc, 0.0, 0.0, st))
            self._all_reduce_small(glob, 2 * Cc)
            cnt = float(R * world)
            self.steps.append(lambda st: cabi.call("gg_bn_bwd_apply", gyp, xp, yp, mp, rp, gp, None, glp, 1, cnt, dxp, None, None,
                                                   R, Cc, act, alpha, st))
        else:
            self.steps.append(lambda st: cabi.call("gg_bn_bwd_apply", gyp, xp, yp, mp, rp, gp, None, dgbp, 1, float(R), dxp, None,
                                                   None, R, Cc, act, alpha, st))

    def _all_reduce_small(self, t, n):
        """in-place sum over ranks of the first n floats of t: one peer-memory kernel inside the graph when available,
        else an eager NCCL all-reduce between graph segments"""
        sar = ggdist.small_all_reduce()
        if sar is not None and n <= sar.MAX_FLOATS:
            fn = lambda st: sar(t, t, n, st)
            fn.is_small_allreduce = True
            self.steps.append(fn)
        else:
            self.steps.append(self._collective(lambda st: ggdist.all_reduce_sum(t)))

    # ---- operations (train ops) ---------------------------------------------------------------
    def _emit_operation(self, op):
        """emit a train op and its scheduling groups: the update itself is a barrier (it waits for everything before it and
        everything after waits for it); the data-parallel gradient buckets in front of it are ordinary groups that depend
        only on the kernels producing their gradients"""
        if op.kind == "group":
            for sub in op.attrs["ops"]:
                self._emit_operation(sub)
            return
        if op.kind == "noop":
            return
        s0 = len(self.steps)
        if op.kind in ("adam", "rmsprop"):
            self._emit_optimizer(op)
            s0 = self._opt_update_start
        elif op.kind == "assign":
            var, val = op.attrs["var"], op.deps[0]
            vp, sp, n = self.rt.param_buffer(var).data_ptr(), self.buf[val.id].data_ptr(), var.size
            self.steps.append(lambda st: cabi.call("gg_unary", cabi.UNARY["copy"], sp, vp, n, 0.0, 0.0, st))
        else:
            raise NotImplementedError("operation %r" % op.kind)
        if len(self.steps) > s0:
            self._add_groups(s0, len(self.steps), set(), "op%d" % op.id, barrier=True)

    def _emit_optimizer(self, op):
        torch = _torch()
        rt = self.rt
        pairs = [(v, g) for v, g in zip(op.attrs["vars"], op.deps) if g is not None]
        params = [rt.param_buffer(v) for v, _ in pairs]
        grads = [self.buf[g.id] for _, g in pairs]
        ms, vs = [], []
        for (v, _), p in zip(pairs, params):
            key = (op.attrs["opt_id"], v.id)
            if key not in rt.slots:
                # Adam: m = v = 0.  RMSProp: tf.train.RMSPropOptimizer creates its `rms` slot as ONES (and `momentum` as
                # zeros); with a zero slot the first updates would be lr*g/sqrt(0.1 g^2) ~ 3.16 lr sign(g) instead of ~lr*g
                first = torch.ones_like(p) if op.kind == "rmsprop" else torch.zeros_like(p)
                rt.slots[key] = (first, torch.zeros_like(p))
                pend = rt.pending_restore.get("slots", {}).pop("%d/%s" % (key[0], v.name), None)   # Saver.restore before run
                if pend is not None:
                    rt.slots[key][0].copy_(pend[0])
                    rt.slots[key][1].copy_(pend[1])
            ms.append(rt.slots[key][0])
            vs.append(rt.slots[key][1])
        if op.attrs["opt_id"] not in rt.opt_state:
            rt.opt_state[op.attrs["opt_id"]] = torch.zeros(3, dtype=torch.float64, device=rt.dev())
            pend = rt.pending_restore.get("opt_state", {}).pop(op.attrs["opt_id"], None)
            if pend is not None:
                rt.opt_state[op.attrs["opt_id"]].copy_(pend)
        state = rt.opt_state[op.attrs["opt_id"]]
        sizes = [v.size for v, _ in pairs]
        world = ggdist.world_size()
        # split update: the variables whose gradients arrive last (`late`) are updated by the barrier launch at the end of
        # the step; all others by an earlier launch that waits only for their own gradients and for the kernels reading them
        late = self._late_vars(pairs) if op.kind == "adam" else set()

        def chunk_table(sel):
            raw = b"".join(struct.pack("<iiq", ti, 0, off) for ti, n in enumerate(sizes) if sel(pairs[ti][0].id)
                           for off in range(0, n, cabi.GG_ADAM_CHUNK))
            return len(raw) // 16, torch.frombuffer(bytearray(raw) or bytearray(16), dtype=torch.uint8).to(rt.dev())
        n_chunks, chk = chunk_table(lambda vid: vid not in late)           # the early set (everything when there is no split)
        n_late, chk_late = chunk_table(lambda vid: vid in late)
        self.keep.append(chk_late)

        def table(gptrs):
            raw = b"".join(struct.pack("<QQQQq", p.data_ptr(), gp, m.data_ptr(), v.data_ptr(), n)
                           for p, gp, m, v, n in zip(params, gptrs, ms, vs, sizes))
            return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(rt.dev())
        tab = table([g.data_ptr() for g in grads])
        self.keep += [chk, tab]
        gscale = 1.0
        adam_tab = tab
        if world > 1:
            # gradients live in (or are copied into) the optimiser's flat buffer; one NCCL all-reduce per readiness bucket
            entries, buckets, total = self.bucket_plan[op.attrs["opt_id"]]
            flat = self.flat[op.attrs["opt_id"]]
            off_of = {e["var"].id: e["off"] for e in entries}
            adam_tab = table([flat.data_ptr() + 4 * off_of[v.id] for v, _ in pairs])
            self.keep += [flat, adam_tab]
            for bi, bucket in enumerate(buckets):
                b0 = len(self.steps)
                reads = set()
                for e in bucket:
                    reads |= self._owners(e["grad"])
                    if not e["direct"]:
                        sp_, dp_, n_ = self.buf[e["grad"].id].data_ptr(), flat.data_ptr() + 4 * e["off"], e["var"].size
                        self.steps.append(lambda st, sp_=sp_, dp_=dp_, n_=n_: cabi.call("gg_unary", cabi.UNARY["copy"], sp_, dp_, n_, 0.0, 0.0, st))
                lo = bucket[0]["off"]
                hi = bucket[-1]["off"] + ((bucket[-1]["var"].size + 63) & ~63)
                view = flat[lo:hi]
                self.keep.append(view)
                fn = self._collective(lambda st, view=view: ggdist.all_reduce_sum(view))
                fn.cost_us = 25.0 + (hi - lo) * 4 / 3.0e5          # ~300 GB/s bus bandwidth + launch latency
                self.steps.append(fn)
                self._add_groups(b0, len(self.steps), reads, "bucket%d_%d" % (op.attrs["opt_id"], bi), barrier=False)
            gscale = 1.0 / world
        tp, cp, sp = adam_tab.data_ptr(), chk.data_ptr(), state.data_ptr()
        a = op.attrs
        if op.kind == "adam" and late:
            lr, b1, b2, eps = a["lr"], a["beta1"], a["beta2"], a["eps"]
            # (1) the state advance: no dependencies, any time before the updates
            t0 = len(self.steps)
            self.steps.append(lambda st: cabi.call("gg_adam_tick", sp, b1, b2, st))
            tick_id = "op%d_tick" % op.id
            self._add_groups(t0, len(self.steps), set(), tick_id, barrier=False)
            # (2) the early update: after the gradients of its variables (their all-reduced buckets under data parallelism)
            # and after the LAST kernel group of every node that reads one of those variables
            early_vars = set(v.id for v, _ in pairs if v.id not in late)
            reads = {tick_id}
            if world > 1:
                for bi, bucket in enumerate(buckets):
                    if any(e["var"].id in early_vars for e in bucket):
                        reads.add("bucket%d_%d" % (op.attrs["opt_id"], bi))
            else:
                for v, g in pairs:
                    if v.id in early_vars:
                        reads |= self._owners(g)
            for gi, grp in enumerate(self.groups):
                if grp["part"][0] == 0 and (grp["reads"] & early_vars):
                    reads.add(self.groups[gi + grp["part"][1] - 1]["writes"])
            e0 = len(self.steps)
            fn = lambda st: cabi.call("gg_adam_apply", tp, cp, n_chunks, sp, lr, b1, b2, eps, gscale, st)
            fn.cost_us = 3.0 + n_chunks * cabi.GG_ADAM_CHUNK * 28 / 5.5e6      # 28 B per parameter at ~5.5 TB/s (L2-resident)
            self.steps.append(fn)
            self._add_groups(e0, len(self.steps), reads, "op%d_early" % op.id, barrier=False)
            # (3) the late update: the step's barrier
            self._opt_update_start = len(self.steps)
            cl = chk_late.data_ptr()
            fn = lambda st: cabi.call("gg_adam_apply", tp, cl, n_late, sp, lr, b1, b2, eps, gscale, st)
            fn.cost_us = 3.0 + n_late * cabi.GG_ADAM_CHUNK * 28 / 5.5e6
            self.steps.append(fn)
            return
        self._opt_update_start = len(self.steps)
        if op.kind == "adam":
            lr, b1, b2, eps = a["lr"], a["beta1"], a["beta2"], a["eps"]
            self.steps.append(lambda st: cabi.call("gg_adam_multi", tp, cp, n_chunks, sp, lr, b1, b2, eps, gscale, st))
        else:
            lr, decay, eps = a["lr"], a["decay"], a["eps"]
            self.steps.append(lambda st: cabi.call("gg_rmsprop_multi", tp, cp, n_chunks, lr, decay, eps, gscale, st))

    # ---- execution ------------------------------------------------------------------------------
    @staticmethod
    def _collective(fn):
        """tag a step as a cross-rank collective: it runs eagerly BETWEEN captured CUDA-graph segments (NCCL stays
        outside stream capture; the kernels on either side of the exchange are still one graph launch each)"""
        fn.is_collective = True
        return fn

    def _launch_all(self):
        st = cabi.stream_ptr()
        for f in self.steps:
            f(st)

    def _schedule(self, n_streams):
        """static list scheduling of the kernel groups onto `n_streams` streams: a group continues the stream of its
        most recent producer when that producer is still the stream's tail, otherwise it takes the least recently used
        stream; cross-stream dependencies become event waits.  Independent branches of the step (E(real) vs G(p_z),
        the two discriminator applications, every wgrad / bias-gradient leaf vs the dgrad chain) then overlap inside the
        captured CUDA graph — most kernels of this workload are latency-bound and fill a fraction of the 148 SMs."""
        return self._schedule_range(range(len(self.groups)), n_streams)

    def _group_cost(self, g):
        """estimated duration (us) of a kernel group inside the graph, for the list scheduler; calibrated on CUPTI
        timelines of the gmgan-CIFAR step (profiles/timeline_*): tensor-core conv / dense launches are dominated by a
        fixed ~10 us of pipeline fill + split-K exchange, element-wise glue by the ~2 us dependent-launch latency"""
        node = g.get("node")
        n_k = g["end"] - g["start"]
        if g.get("collective"):
            return float(getattr(self.steps[g["start"]], "cost_us", 60.0))
        if node is None:                                   # optimiser update (whole / early / late part) / rng tick
            return float(getattr(self.steps[g["start"]], "cost_us", 16.0 if g["barrier"] else 2.0))
        mb = node.size * 4 / 1e6
        op = node.op
        if op == "conv":
            a = node.attrs
            gf = 2.0 * a["B"] * a["Ho"] * a["Wo"] * a["Co"] * a["Ci"] * a["k"] * a["k"] / 1e9
            if a["Ci"] <= 4 and a["mode"] in ("fwd", "dgrad"):
                return 5.0 + 20.0 * gf                      # one CUDA-core launch (gg_conv_small.cu)
            t = 10.0 + 3.0 * gf
            if a["Ci"] <= 4:
                t += 10.0                                   # patch-matrix kernel in front of the wgrad GEMM
            return t
        if op == "matmul":
            M, N = node.shape
            K = node.inputs[0].shape[0] if node.attrs["ta"] else node.inputs[0].shape[1]
            return 12.0 if (N % 32 == 0 and K % 32 == 0) else 8.0
        if op == "bn":
            return 4.0 + 1.5 * mb
        if op == "bn_grad":
            return 5.0 + 3.0 * mb
        return 2.0 * n_k + 0.5 * mb

    def _schedule_range(self, idxs, n_streams):
        """Static list scheduling (HEFT-style) of the kernel groups `idxs` onto n_streams streams of one CUDA graph.

        Groups become ready when all their producers are placed; among the ready ones the group with the longest
        remaining dependency chain (bottom level, from the cost model above) goes first, onto the stream where it can
        start earliest in a simulated timeline — preferring the stream of the producer it waits for, so that chains stay
        on one stream and cross-stream edges (event waits) appear only where branches fork or join.  A stream runs its
        groups in issue order, so a bad placement is a false dependency: the previous policy (continue the producer's
        stream, else least-recently-used) left the Extractor's backward chain queued behind the Generator's although
        the two are independent (CUPTI timeline, profiles/timeline_gen_r1_before.txt).
        Returns (issue order, stream of each group, cross-stream waits, groups that need an event)."""
        idxs = list(idxs)
        pos = {gi: k for k, gi in enumerate(idxs)}
        cost = {gi: self._group_cost(self.groups[gi]) for gi in idxs}
        # dependencies: producers of what the group reads; an optimiser step (barrier) waits for everything before it and
        # everything after it waits for the barrier
        producer, deps, last_barrier, last_ordered, seen = {}, {}, None, {}, []
        for gi in idxs:
            g = self.groups[gi]
            d = set(producer[o] for o in g["reads"] if o in producer)
            if last_barrier is not None:
                d.add(last_barrier)
            if g.get("ordered"):
                if g["ordered"] in last_ordered:
                    d.add(last_ordered[g["ordered"]])
                last_ordered[g["ordered"]] = gi
            if g["barrier"]:
                d |= set(seen)
                last_barrier = gi
            deps[gi] = d
            producer[g["writes"]] = gi
            seen.append(gi)
        succ = {gi: [] for gi in idxs}
        for gi in idxs:
            for d in deps[gi]:
                succ[d].append(gi)
        bottom = {}
        for gi in reversed(idxs):
            bottom[gi] = cost[gi] + max([bottom[c] for c in succ[gi]] or [0.0])
        if n_streams <= 1 or os.environ.get("GG_SCHED", "heft") == "order":
            return self._schedule_in_order(idxs, deps, n_streams)
        import heapq
        indeg = {gi: len(deps[gi]) for gi in idxs}
        # Issue order.  "heft": longest remaining chain first — the critical chain gets its streams, but every group with a
        # short tail (weight gradients, bias-gradient reductions: only the update waits for them) is issued LAST, i.e. queued
        # behind everything its stream already holds although it was ready long before (profiles/timeline_disc_r2.txt: ~40
        # glue kernels after the last backward kernel).  "asap": earliest possible start first (ties: longest chain), so
        # that such groups fill the idle streams while the chain runs.
        mode = os.environ.get("GG_SCHED", "asap")
        est = {}
        for gi in idxs:
            est[gi] = max([est[d] + cost[d] for d in deps[gi]] or [0.0])

        def prio(gi):
            return (est[gi], -bottom[gi], pos[gi], gi) if mode == "asap" else (-bottom[gi], pos[gi], 0, gi)
        ready = [prio(gi) for gi in idxs if indeg[gi] == 0]
        heapq.heapify(ready)
        free_at = [0.0] * n_streams
        finish, assign, waits, need_event, order, issued = {}, {}, {}, set(), [], {}
        sync_cost = 1.0                       # a cross-stream edge costs an event wait
        # Streams are FIFOs and list scheduling cannot back-fill: a low-priority group (a gradient bucket's all-reduce has
        # only the update behind it) is issued late and would queue behind whatever its stream already holds — the round-2
        # timeline showed both bucket all-reduces starting after the LAST backward kernel.  Collectives therefore get a
        # stream of their own: in the captured graph they depend on exactly their producers.
        has_coll = any(self.groups[gi]["collective"] for gi in idxs)
        comm_stream = n_streams - 1 if (has_coll and n_streams > 2) else None
        # Stream classes (GG_PRIO=1, >= 4 streams): groups that can slip by `slack` without moving the end of the step — weight
        # gradients, bias-gradient reductions — go to LOW-priority streams, everything near the critical chain to HIGH-priority
        # ones (_stream_priorities): a 128-CTA weight-gradient launch then no longer holds the SMs the chain's next launch
        # needs; the block scheduler hands freed SMs to the pending high-priority grid first.
        prio_on = os.environ.get("GG_PRIO", "1") == "1" and n_streams >= 4
        pr = self._stream_priorities(n_streams, comm_stream is not None) if prio_on else None
        makespan = max([est[gi] + bottom[gi] for gi in idxs] or [0.0])
        slack_min = float(os.environ.get("GG_PRIO_SLACK_US", "40"))
        low_cls = {gi: (makespan - (est[gi] + bottom[gi]) > slack_min) for gi in idxs}
        # ... but never a group that a HIGH-class group waits for: the CUPTI timeline of the G step (profiles/timeline_gen_r2.txt)
        # has a 40 us hole with NO kernel in flight right after fake_x is complete — the image decode / dequantisation chain the
        # discriminator's first layer also needs had been ready since t = 10 us, but sat on a low-priority stream (its modelled
        # slack is 75 us) and low-priority queues are not served while a high-priority queue holds pending work, even work that
        # is itself blocked on them: a priority inversion.  Low class = slack AND only the update (or other low groups) behind it.
        # Measured (gpurun_out/quick_s27.txt, quick_s28.txt; ms per iteration, off / on / cheap-glue-only): face 1.844 / 1.795 / 1.826,
        # SSGAN 6.31 / 5.69 / 5.86 — but gmgan-CIFAR 0.797 / 0.848 / 0.814: at batch 64 no launch fills the GPU and the step is
        # its latency chain, which any early co-runner stretches.  "auto" turns the rule on for plans whose tensor-core launches are
        # multi-wave (more output tiles than SMs), i.e. throughput-bound steps.
        tail_mode = os.environ.get("GG_PRIO_TAIL", "auto")
        if tail_mode == "auto":
            convs = [self.groups[gi]["node"] for gi in idxs if self.groups[gi].get("node") is not None and self.groups[gi]["node"].op == "conv"]
            big = sum(1 for n in convs if self._conv_tiles(n) > 148)
            tail_mode = "1" if convs and big * 4 >= len(convs) else "0"
        self.prio_tail_mode = tail_mode
        if tail_mode in ("1", "2"):
            for gi in reversed(idxs):
                if low_cls[gi] and any(not self.groups[c]["barrier"] and not low_cls[c] for c in succ[gi]):
                    if tail_mode == "2" and cost[gi] >= 8.0:
                        continue                      # mode 2: only the cheap glue is pulled up, heavy launches keep their class
                    low_cls[gi] = False
        self.low_class = low_cls
        while ready:
            gi = heapq.heappop(ready)[-1]
            best = None
            is_coll = self.groups[gi]["collective"]
            for s in range(n_streams):
                if comm_stream is not None and (s == comm_stream) != bool(is_coll):
                    continue
                if pr is not None and not is_coll and (pr[s] == 0) != low_cls[gi]:
                    continue
                t = free_at[s]
                for d in deps[gi]:
                    t = max(t, finish[d] + (0.0 if assign[d] == s else sync_cost))
                key = (t, 0 if any(assign[d] == s for d in deps[gi]) else 1, s)
                if best is None or key < best[0]:
                    best = (key, s, t)
            _, s, t = best
            assign[gi] = s
            finish[gi] = t + cost[gi]
            free_at[s] = finish[gi]
            # only the latest cross-stream producers per foreign stream need a wait (stream order covers the earlier ones)
            per_stream = {}
            for d in deps[gi]:
                q = assign[d]
                if q != s and (q not in per_stream or issued[d] > issued[per_stream[q]]):
                    per_stream[q] = d
            w = sorted(per_stream.values(), key=lambda d: issued[d])
            waits[gi] = w
            need_event.update(w)
            issued[gi] = len(order)
            order.append(gi)
            for c in succ[gi]:
                indeg[c] -= 1
                if indeg[c] == 0:
                    heapq.heappush(ready, prio(c))
        assert len(order) == len(idxs)
        self.sched_estimate_us = max(finish.values()) if finish else 0.0
        return order, assign, waits, need_event

    @staticmethod
    def _conv_tiles(node):
        """128 x 128 output tiles of a conv node's implicit GEMM (fwd / dgrad: pixels x channels; wgrad: filter rows x channels)"""
        a = node.attrs
        if a["mode"] == "wgrad":
            return -(-a["k"] * a["k"] * a["Ci"] // 128) * -(-a["Co"] // 128)
        if a["mode"] == "fwd":
            return -(-a["B"] * a["Ho"] * a["Wo"] // 128) * -(-a["Co"] // 128)
        return -(-a["B"] * a["H"] * a["W"] // 128) * -(-a["Ci"] // 128)

    @staticmethod
    def _stream_priorities(n_streams, has_comm):
        """priority of each stream of a captured step: 0 = low (the capture's own stream, index 0, and the last side streams),
        -1 = high (side streams 1..h and, under data parallelism, the collectives' stream)"""
        n_side = n_streams - 1 - (1 if has_comm else 0)
        n_low_side = max(1, n_side // 3)
        return [0] + [-1] * (n_side - n_low_side) + [0] * n_low_side + ([-1] if has_comm else [])

    def _schedule_in_order(self, idxs, deps, n_streams):
        """the plan's own topological order; a group continues the stream of its most recent producer when that producer is
        still the stream's tail, otherwise it takes the least recently used stream (GG_SCHED=order, and n_streams == 1)"""
        assign, waits, need_event = {}, {}, set()
        tail = [None] * n_streams
        for gi in idxs:
            d = deps[gi]
            cand = [s for s in range(n_streams) if tail[s] is not None and tail[s] in d]
            if cand:
                s = max(cand, key=lambda q: tail[q])
            else:
                free = [q for q in range(n_streams) if tail[q] is None]
                s = free[0] if free else min(range(n_streams), key=lambda q: tail[q])
            w = sorted(x for x in d if assign[x] != s)
            need_event.update(w)
            assign[gi] = s
            waits[gi] = w
            tail[s] = gi
        return list(idxs), assign, waits, need_event

    def _capture_range(self, idxs, n_streams):
        """capture the kernel groups `idxs` (no collectives among them) into one CUDA graph, spread over n_streams streams"""
        torch = _torch()
        order, assign, waits, need_event = self._schedule_range(idxs, n_streams)
        g = torch.cuda.CUDAGraph()
        has_comm = any(self.groups[gi]["collective"] for gi in idxs) and n_streams > 2
        if os.environ.get("GG_PRIO", "1") == "1" and n_streams >= 4:
            pr = self._stream_priorities(n_streams, has_comm)
            side = [torch.cuda.Stream(priority=pr[i + 1]) for i in range(n_streams - 1)]   # kernel nodes inherit the priority
        else:
            side = [torch.cuda.Stream() for _ in range(max(n_streams - 1, 0))]
        self.keep.append(side)
        trace = os.environ.get("GG_TRACE", "0") == "1"      # tools/trace_step.py: timed event nodes around every group
        with torch.cuda.graph(g):
            cs = torch.cuda.current_stream()
            streams = [cs] + side
            if trace:
                self.trace_t0 = self._trace_event(cs.cuda_stream)
            for sd in side:
                sd.wait_stream(cs)
            events = {}
            for gi in order:
                grp = self.groups[gi]
                st = streams[assign[gi]]
                for d in waits[gi]:
                    st.wait_event(events[d])
                sp = st.cuda_stream
                if trace:
                    ta = self._trace_event(sp)
                if grp["collective"]:
                    with torch.cuda.stream(st):        # NCCL (torch.distributed) issues on torch's current stream
                        for f in self.steps[grp["start"]:grp["end"]]:
                            f(sp)
                else:
                    for f in self.steps[grp["start"]:grp["end"]]:
                        f(sp)
                if trace:
                    self.trace.append((gi, assign[gi], ta, self._trace_event(sp), list(waits[gi])))
                if gi in need_event:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    events[gi] = ev
            for sd in side:
                cs.wait_stream(sd)
            self.keep.append(events)
        return g

    @staticmethod
    def _trace_event(stream_ptr):
        ev = C.c_void_p()
        cabi.call("gg_trace_event_create", C.byref(ev))
        cabi.call("gg_trace_event_record", ev, stream_ptr)
        return ev

    @staticmethod
    def trace_elapsed_us(a, b):
        us = C.c_float()
        cabi.call("gg_trace_event_elapsed_us", a, b, C.byref(us))
        return us.value

    def _capture_segments(self):
        """launch list -> [CUDA graph | eager collective | CUDA graph | ...]: NCCL collectives stay outside stream capture
        and run on the main stream between graph segments; every segment is scheduled over n_streams streams"""
        torch = _torch()
        segments, cur = [], []
        # GG_NCCL_IN_GRAPH=1: capture the NCCL all-reduce as a node of the step's graph (one graph launch per step, and the
        # scheduler can run independent kernels beside the exchange) instead of cutting the launch list around it
        in_graph = os.environ.get("GG_NCCL_IN_GRAPH", "1") == "1"
        for gi, grp in enumerate(self.groups):
            if grp["collective"] and not in_graph:
                if cur:
                    segments.append(cur)
                    cur = []
                segments.append(self.steps[grp["start"]])
            else:
                cur.append(gi)
        if cur:
            segments.append(cur)
        torch.cuda.synchronize()
        before = cabi.lib.gg_launch_count()
        out = [seg if callable(seg) else self._capture_range(seg, self.n_streams) for seg in segments]
        self.kernel_launches = cabi.lib.gg_launch_count() - before
        return out

    def run(self, feed_dict, to_host=True, deferred=False):
        torch = _torch()
        for node, value in feed_dict.items():
            d, _h = self.rt.feed_buffer(node)
            if isinstance(value, torch.Tensor):
                d.copy_(value.reshape(-1), non_blocking=True)   # already-resident input (kernel-only timing)
            else:
                arr = np.asarray(value)
                if arr.size != node.size:
                    raise ValueError("feed for %s has %d elements, expected %s" % (node.name, arr.size, tuple(node.shape)))
                if arr.dtype == np.uint8 and node.dtype == int32:
                    d8, _h8 = self.rt.feed_buffer_u8(node)
                    slot = self.rt.stage_slot(("h2d8", node.id), node.size, torch.uint8)
                    slot[0].copy_(torch.from_numpy(np.ascontiguousarray(arr.reshape(-1))))
                    d8.copy_(slot[0], non_blocking=True)
                    slot[1] = torch.cuda.Event()
                    slot[1].record()
                    cabi.call("gg_widen_u8_i32", d8.data_ptr(), d.data_ptr(), node.size, cabi.stream_ptr())
                    self.h2d_bytes = getattr(self, "h2d_bytes", 0) + arr.size
                    continue
                self.h2d_bytes = getattr(self, "h2d_bytes", 0) + arr.size * 4
                slot = self.rt.stage_slot(("h2d", node.id), node.size, d.dtype)
                slot[0].copy_(torch.from_numpy(np.ascontiguousarray(arr.reshape(-1).astype(node.dtype.as_numpy_dtype, copy=False))))
                d.copy_(slot[0], non_blocking=True)
                slot[1] = torch.cuda.Event()
                slot[1].record()
        self.runs += 1
        if self.rt.use_cuda_graph:
            if self.graph is None:
                # capture without executing: no side effect (Adam / RNG tick) happens until the first replay
                self.graph = self._capture_segments()
            st = None
            for seg in self.graph:
                if callable(seg):
                    st = st or cabi.stream_ptr()
                    seg(st)
                else:
                    seg.replay()
        else:
            before = cabi.lib.gg_launch_count()
            self._launch_all()
            self.kernel_launches = cabi.lib.gg_launch_count() - before
        if not to_host:
            return None
        # device -> host: every fetched tensor is copied into a pinned slot by THIS run; the host waits for the copy's event
        # here (TensorFlow's semantics) or, with deferred fetches, when the value is first used
        out = []
        for f in self.fetches:
            if isinstance(f, Tensor):
                t = self.buf[f.id][:max(f.size, 1)]
                slot = self.rt.stage_slot(("d2h", id(self), f.id), t.numel(), t.dtype)
                slot[0].copy_(t, non_blocking=True)
                slot[1] = torch.cuda.Event()
                slot[1].record()
                self.d2h_bytes = getattr(self, "d2h_bytes", 0) + t.numel() * t.element_size()
                val = Deferred(slot[0], slot[1], f.shape)
                if deferred:
                    slot[2] = val
                out.append(val)
            else:
                out.append(None)
        if not deferred:
            out = [v.result() if v is not None else None for v in out]
        return out

    def launches_per_run(self):
        return len(self.steps)
