"""Element-wise cluster fusion for the plan compiler (gg/executor.py).

The reference's script-level glue is long chains of tiny tf.* ops — the Gumbel noise `-tf.log(-tf.log(U + eps) + eps)` and
the soft assignment around it (gmgan_inference_cifar10.py:155-163), the image decode `2*((tf.cast(x)/255.)-.5)` (:341-342),
the mixture-prior distances, the sigmoid-cross-entropy terms of tflib/objs/gan_inference.py:85-101 and all their gradients.
One launch per op is 2-3 us of latency on a KB-sized tensor, ~110 of the ~285 launches of a gmgan-CIFAR iteration.  This
module groups connected element-wise nodes of a plan into clusters; each cluster becomes ONE `gg_ew_run` launch of a small
register program (include/gg_b200.h: gg_ew_program) that evaluates the whole group per output element, optionally followed by
the row reduction that consumes it (`tf.reduce_mean` of the per-example losses).  The arithmetic per op is the code the one-op
kernels run, in the same order, so a fused plan is bit-identical to the unfused one (GG_FUSE_EW=0).

Vocabulary: a cluster iterates over D = the non-1 extents of its largest member ("iteration space"); every member and every
external input spans a subset of those extents (its `mask`) and is broadcast along the others, which is how
`tf.expand_dims` / implicit broadcasting inside a chain is evaluated without materialising the broadcast.
"""
from . import cabi
from .graph import float32, int32, prod

EW_OPS = ("unary", "binary", "broadcast", "add_n", "cast")
TRANSPARENT = ("reshape", "stop_gradient")
MAX_ELEMS_BROADCAST_MERGE = 1 << 22     # a smaller-shaped producer is re-evaluated per element of the consumer's space
REDUCE_CODE = {"sum": 1, "mean": 2, "max": 3}


def squeeze(shape):
    return tuple(int(s) for s in shape if s != 1)


class Cluster(object):
    def __init__(self, members):
        self.members = members          # element-wise nodes, topological order
        self.aliases = []               # transparent reshape / stop_gradient nodes whose source is a member
        self.reduce = None              # the row reduction folded behind the last member
        self.key = None                 # scheduling id (owner) of everything the launch writes
        self.desc = None                # the program, as plain Python (tests evaluate it with numpy)
        self.emitted = False

    def nodes(self):
        return self.members + self.aliases + ([self.reduce] if self.reduce is not None else [])


class Planner(object):
    """clusters for one plan: `order` (topological), ids of fed nodes, nodes that must own a buffer for reasons the graph
    does not show (fetches, optimiser inputs, activation masks of fused dgrad launches), nodes computed elsewhere"""

    def __init__(self, order, fed, external_use, excluded, extra_edges=()):
        self.order = order
        self.fed = fed
        self.pos = {n.id: i for i, n in enumerate(order)}
        self.node = {n.id: n for n in order}
        self.excluded = excluded
        self.consumers = {n.id: [] for n in order}
        for n in order:
            if n.id in fed:
                continue
            for i in n.inputs:
                if i.id in self.consumers:
                    self.consumers[i.id].append(n)
        for src, dst in extra_edges:     # operands a peephole added to a launch (activation masks, gather addends): real reads
            if src.id in self.consumers and dst.id in self.pos:
                self.consumers[src.id].append(dst)
        self.ext = external_use          # id -> count
        self.parent = {}
        self.members_of = {}

    # ---- candidates and edges -------------------------------------------------------------------------------------
    def candidate(self, n):
        if n.op not in EW_OPS or n.id in self.fed or n.id in self.excluded or n.dtype != float32 or n.size <= 0:
            return False
        if n.op == "cast":
            return n.inputs[0].dtype == int32
        return all(i.dtype == float32 for i in n.inputs) and len(n.inputs) >= 1

    def source(self, t):
        """the node whose storage `t` aliases through transparent (extent-preserving) reshapes, and the alias chain"""
        chain = []
        while t.op in TRANSPARENT and t.id not in self.fed and t.id in self.pos and squeeze(t.shape) == squeeze(t.inputs[0].shape):
            chain.append(t)
            t = t.inputs[0]
        return t, chain

    def find(self, i):
        while self.parent.get(i, i) != i:
            self.parent[i] = self.parent.get(self.parent[i], self.parent[i])
            i = self.parent[i]
        return i

    def cluster_ids(self, i):
        return self.members_of.get(self.find(i), [i])

    def all_consumers(self, nid):
        """consumers of a node, looking through transparent aliases (which carry its storage on)"""
        out, stack = [], list(self.consumers.get(nid, ()))
        while stack:
            c = stack.pop()
            if c.op in TRANSPARENT and c.id not in self.fed and squeeze(c.shape) == squeeze(c.inputs[0].shape):
                if self.ext.get(c.id, 0):
                    out.append(None)                    # the alias itself is needed outside: storage must exist
                stack.extend(self.consumers.get(c.id, ()))
            else:
                out.append(c)
        return out

    def needs_buffer(self, nid, inside):
        if self.ext.get(nid, 0):
            return True
        return any(c is None or c.id not in inside for c in self.all_consumers(nid))

    def reaches(self, a_ids, b_ids):
        """is there a path from cluster A to cluster B through at least one node outside both (cluster-level hops)?"""
        a_set, b_set = set(a_ids), set(b_ids)
        limit = max(self.pos[i] for i in b_ids)
        seen_nodes, seen_clusters = set(), set()
        stack = []
        for i in a_ids:
            for c in self.consumers.get(i, ()):
                stack.append((c, True))
        while stack:
            n, direct = stack.pop()
            if n.id in a_set:
                continue
            if n.id in b_set:
                if not direct:
                    return True
                continue
            if (n.id, direct) in seen_nodes:
                continue
            seen_nodes.add((n.id, direct))
            root = self.find(n.id)
            if root in self.members_of:
                if root in seen_clusters:
                    continue
                seen_clusters.add(root)
                for m in self.members_of[root]:
                    for c in self.consumers.get(m, ()):
                        stack.append((c, False))
                continue
            if self.pos[n.id] > limit:
                continue
            transparent = n.op in TRANSPARENT and n.id not in self.fed and squeeze(n.shape) == squeeze(n.inputs[0].shape)
            for c in self.consumers.get(n.id, ()):
                stack.append((c, direct and transparent))
        return False

    # ---- clustering -----------------------------------------------------------------------------------------------
    def size_ok(self, ids):
        inside = set(ids)
        n_instr, loads, outs = 0, set(), 0
        for i in ids:
            n = self.node[i]
            if n.op in ("unary", "binary"):
                n_instr += 1
            elif n.op == "add_n":
                n_instr += len(n.inputs) - 1
            for inp in n.inputs:
                src, _ = self.source(inp)
                if src.id not in inside:
                    loads.add(src.id)
            if self.needs_buffer(i, inside):
                outs += 1
        # loads of one tensor under two different broadcast masks count twice: leave headroom
        return n_instr <= cabi.GG_EW_MAX_INSTR and len(loads) <= cabi.GG_EW_MAX_IN - 2 and outs <= cabi.GG_EW_MAX_OUT \
            and len(loads) + n_instr <= cabi.GG_EW_REGS + 8

    def space(self, ids):
        return max((squeeze(self.node[i].shape) for i in ids), key=lambda s: (prod(s), len(s)))

    def build(self):
        # consumers first: when a producer is looked at, the cluster of each of its consumers is already as large as it will
        # get on that side, so a small-shaped producer (a scalar `cost/B`, a [K] row of log-priors) is re-evaluated inside the
        # ONE cluster that consumes it instead of being glued to its siblings through a shared scalar upstream
        for v in reversed(self.order):
            if not self.candidate(v):
                continue
            for inp in v.inputs:
                src, _ = self.source(inp)
                if src.id not in self.pos or not self.candidate(src):
                    continue
                ra, rb = self.find(src.id), self.find(v.id)
                if ra == rb:
                    continue
                a_ids, b_ids = self.cluster_ids(src.id), self.cluster_ids(v.id)
                merged = a_ids + b_ids
                sa, sb = self.space(a_ids), self.space(b_ids)
                if sa != sb:
                    small, big = (a_ids, b_ids) if prod(sa) <= prod(sb) else (b_ids, a_ids)
                    if big is a_ids:                         # a consumer is never smaller than its producer's space
                        continue
                    inside = set(merged)
                    if prod(self.space(big)) > MAX_ELEMS_BROADCAST_MERGE or any(self.needs_buffer(i, inside) for i in small):
                        continue
                if not self.size_ok(merged):
                    continue
                if self.reaches(a_ids, b_ids) or self.reaches(b_ids, a_ids):
                    continue
                # trial program: broadcast alignment, register pressure and the operand limits are checked exactly
                if not self._program(Cluster([self.node[i] for i in sorted(merged, key=lambda i: self.pos[i])])):
                    continue
                self.parent[ra] = rb
                self.members_of.pop(ra, None)
                self.members_of[rb] = sorted(merged, key=lambda i: self.pos[i])
        clusters = []
        for root, ids in self.members_of.items():
            if self.find(root) != root:
                continue
            cl = Cluster([self.node[i] for i in ids])
            self._attach_reduce(cl)
            work = len(cl.members)                       # every member is a launch of the unfused plan
            if work + (1 if cl.reduce is not None else 0) < 2:
                continue                                 # nothing to save: the specialised one-op kernel stays
            if not self._program(cl):
                if cl.reduce is None:
                    continue
                cl.reduce = None
                if work < 2 or not self._program(cl):
                    continue
            clusters.append(cl)
        # single element-wise nodes feeding a row reduction (the BCE -> mean of every objective): 2 launches -> 1
        taken = set(n.id for cl in clusters for n in cl.nodes())
        for v in self.order:
            if v.id in taken or not self.candidate(v) or v.op not in ("unary", "binary"):
                continue
            cl = Cluster([v])
            self._attach_reduce(cl)
            if cl.reduce is not None and cl.reduce.id not in taken and self._program(cl):
                clusters.append(cl)
                taken.update(n.id for n in cl.nodes())
        return clusters

    def _attach_reduce(self, cl):
        root = cl.members[-1]
        inside = set(n.id for n in cl.members)
        if self.ext.get(root.id, 0) or squeeze(root.shape) != self.space(list(inside)):
            return
        cons = self.consumers.get(root.id, ())
        if len(cons) != 1:
            return
        z = cons[0]
        if z.op != "reduce" or z.id in self.fed or z.id in self.excluded or z.inputs[0] is not root or z.attrs["fn"] not in REDUCE_CODE:
            return
        axes = z.attrs["axes"]
        shp = root.shape
        if prod(shp[axes[-1] + 1:]) != 1 or prod(shp[axes[0]:axes[-1] + 1]) < 1:
            return
        # the reduced launch runs one CTA per row — exactly what gg_reduce does whenever inner == 1, so nothing is serialised
        cl.reduce = z

    # ---- program generation ---------------------------------------------------------------------------------------
    def _program(self, cl):
        members = cl.members
        inside = set(n.id for n in members)
        D = self.space(list(inside))
        nd = len(D)
        full = tuple([True] * nd)
        mask = {}
        loads = {}                       # (source id, mask) -> dict(node=direct input, src=source)
        operand = {}                     # (member id, slot) -> ("m", source id) | ("l", load key)
        for v in reversed(members):
            if v.id not in mask:
                if squeeze(v.shape) != D:
                    return False
                mask[v.id] = full
            mv = mask[v.id]
            dpos = [p for p in range(nd) if mv[p]]
            nz = [i for i, e in enumerate(v.shape) if e != 1]
            if len(nz) != len(dpos) or any(v.shape[i] != D[p] for i, p in zip(nz, dpos)):
                return False
            where = dict(zip(nz, dpos))
            for slot, top in enumerate(v.inputs):
                mt = [False] * nd
                shift = len(v.shape) - len(top.shape)
                if shift < 0:
                    return False
                for j, e in enumerate(top.shape):
                    if e == 1:
                        continue
                    i = j + shift
                    if i not in where or v.shape[i] != e:
                        return False
                    mt[where[i]] = True
                mt = tuple(mt)
                src, chain = self.source(top)
                if src.id in inside:
                    if mask.setdefault(src.id, mt) != mt:
                        return False
                    operand[(v.id, slot)] = ("m", src.id)
                    for r in chain:
                        if r not in cl.aliases:
                            cl.aliases.append(r)
                else:
                    key = (src.id, mt)
                    loads.setdefault(key, dict(node=top, src=src, mask=mt, is_int=(top.dtype == int32)))
                    operand[(v.id, slot)] = ("l", key)
        if len(loads) > cabi.GG_EW_MAX_IN:
            return False
        # transparent aliases of members consumed OUTSIDE the cluster also hang off it (they alias a member's buffer)
        for m in members:
            stack = [c for c in self.consumers.get(m.id, ())]
            while stack:
                c = stack.pop()
                if c.op in TRANSPARENT and c.id not in self.fed and squeeze(c.shape) == squeeze(c.inputs[0].shape):
                    if c not in cl.aliases:
                        cl.aliases.append(c)
                    stack.extend(self.consumers.get(c.id, ()))
        cl.aliases.sort(key=lambda n: self.pos[n.id])
        all_inside = inside | set(a.id for a in cl.aliases)
        # outputs: members some node outside the cluster reads (directly or through an alias)
        outputs = []
        for m in members:
            need = bool(self.ext.get(m.id, 0))
            for c in self.all_consumers(m.id):
                if c is None or (c.id not in inside and not (cl.reduce is not None and c is cl.reduce and m is members[-1])):
                    need = True
            if need:
                if mask[m.id] != full:
                    return False
                outputs.append(m)
        red = None
        if cl.reduce is not None:
            root = members[-1]
            if root in outputs:
                return False
            z = cl.reduce
            axes = z.attrs["axes"]
            n_red = len(squeeze(root.shape[axes[0]:axes[-1] + 1]))
            red = dict(node=z, n_red_dims=n_red, red=prod(root.shape[axes[0]:axes[-1] + 1]), op=REDUCE_CODE[z.attrs["fn"]])
            if n_red == 0:
                return False                     # reducing extent-1 axes only: a copy, leave it alone
        elif not outputs:
            return False
        if len(outputs) + (1 if red else 0) > cabi.GG_EW_MAX_OUT:
            return False
        # iteration dims: merge neighbouring extents on which every load agrees (never across the reduction boundary)
        load_list = list(loads.values())
        groups = []
        boundary = nd - red["n_red_dims"] if red else None
        for p in range(nd):
            if groups and p != boundary and all(ld["mask"][p] == ld["mask"][p - 1] for ld in load_list):
                groups[-1].append(p)
            else:
                groups.append([p])
        if red:
            inner_groups = [g for g in groups if g[0] >= boundary]
            if len(inner_groups) != 1:
                return False
        if len(groups) > 4:
            return False
        dims = [prod(D[p] for p in g) for g in groups]
        if red and not [g for g in groups if g[0] < boundary]:
            groups, dims = [[]] + groups, [1] + dims
        while len(dims) < 4:
            groups, dims = [[]] + groups, [1] + dims
        if red:
            # kernel convention: the reduced extent is dims[3]
            assert dims[3] == red["red"], (dims, red["red"])
        for ld in load_list:
            strides = []
            for g in groups:
                if not g or not ld["mask"][g[-1]]:
                    strides.append(0)
                else:
                    strides.append(prod(D[q] for q in range(g[-1] + 1, nd) if ld["mask"][q]))
            ld["strides"] = strides
        flat = all(ld["mask"] == full for ld in load_list)
        # instructions + register allocation (linear scan; registers 0..n_in-1 hold the loads)
        value_of = {}                                  # member id -> value id
        instrs = []                                    # dict(kind, op, dst(value), src0(value), src1(value), a, b)
        n_in = len(load_list)
        load_value = {key: k for k, key in enumerate(loads.keys())}
        next_value = [n_in]

        def val(v, slot):
            kind, ref = operand[(v.id, slot)]
            return value_of[ref] if kind == "m" else load_value[ref]

        def new_value():
            next_value[0] += 1
            return next_value[0] - 1
        for v in members:
            if v.op in ("broadcast", "cast"):
                value_of[v.id] = val(v, 0)
            elif v.op == "unary":
                d = new_value()
                instrs.append(dict(kind=0, op=cabi.UNARY[v.attrs["fn"]], dst=d, src0=val(v, 0), src1=0, a=float(v.attrs["a"]), b=float(v.attrs["b"])))
                value_of[v.id] = d
            elif v.op == "binary":
                d = new_value()
                instrs.append(dict(kind=1, op=cabi.BINARY[v.attrs["fn"]], dst=d, src0=val(v, 0), src1=val(v, 1), a=float(v.attrs["alpha"]), b=0.0))
                value_of[v.id] = d
            else:                                      # add_n: ((x0 + x1) + x2) + ... like gg_add_n
                acc = val(v, 0)
                for slot in range(1, len(v.inputs)):
                    d = new_value()
                    instrs.append(dict(kind=1, op=cabi.BINARY["add"], dst=d, src0=acc, src1=val(v, slot), a=0.0, b=0.0))
                    acc = d
                value_of[v.id] = acc
        if len(instrs) > cabi.GG_EW_MAX_INSTR:
            return False
        out_nodes = ([members[-1]] if red else []) + outputs
        out_values = [value_of[m.id] for m in out_nodes]
        last_use = {}
        for j, q in enumerate(instrs):
            last_use[q["src0"]] = j
            if q["kind"] == 1:
                last_use[q["src1"]] = j
        for ov in out_values:
            last_use[ov] = len(instrs) + 1
        reg_of = {k: k for k in range(n_in)}
        free = list(range(cabi.GG_EW_REGS - 1, n_in - 1, -1))
        for j, q in enumerate(instrs):
            srcs = [q["src0"]] + ([q["src1"]] if q["kind"] == 1 else [])
            regs = [reg_of[s] for s in srcs]
            for s in set(srcs):
                if last_use.get(s, -1) == j:
                    free.append(reg_of[s])               # the kernel reads its sources before it writes dst
            if not free:
                return False
            r = free.pop()
            reg_of[q["dst"]] = r
            if q["dst"] not in last_use:
                free.append(r)                           # dead value (cannot happen for a pruned graph)
            q["rdst"], q["rsrc0"], q["rsrc1"] = r, regs[0], (regs[1] if len(regs) > 1 else 0)
        cl.desc = dict(D=D, dims=dims, flat=flat, loads=load_list, instrs=instrs, out_nodes=out_nodes,
                       out_regs=[reg_of[v] for v in out_values], reduce=red, interior=[m for m in members if m not in outputs])
        cl.key = (cl.reduce if cl.reduce is not None else members[-1]).id
        return True


def to_struct(desc, in_ptrs, out_ptrs):
    """the ctypes gg_ew_program of a cluster description, with device pointers filled in"""
    p = cabi.EwProgram()
    p.n_in, p.n_out, p.n_instr = len(desc["loads"]), len(out_ptrs), len(desc["instrs"])
    p.flat = 1 if desc["flat"] else 0
    p.reduce_op = desc["reduce"]["op"] if desc["reduce"] else 0
    for i in range(4):
        p.dims[i] = desc["dims"][i]
    for k, ld in enumerate(desc["loads"]):
        p.inp[k] = in_ptrs[k]
        p.in_is_int[k] = 1 if ld["is_int"] else 0
        for i in range(4):
            p.in_stride[k][i] = ld["strides"][i]
    for k, ptr in enumerate(out_ptrs):
        p.out[k] = ptr
        p.out_reg[k] = desc["out_regs"][k]
    for j, q in enumerate(desc["instrs"]):
        ins = p.instr[j]
        ins.kind, ins.op, ins.dst, ins.src0, ins.src1, ins.a, ins.b = q["kind"], q["op"], q["rdst"], q["rsrc0"], q["rsrc1"], q["a"], q["b"]
    return p
