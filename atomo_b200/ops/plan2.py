"""Planner of the overlapped / sharded bf16 engine (``runtime/shadow_engine.py``, kernels ``csrc/v2_*.cu``).

What changes against ``ops/plan.py`` (the fp32-flat plan of round 1):

* **Where the bytes live.**  Conv / linear weights are bf16 *leaf* tensors in the memory format cuDNN wants
  (``channels_last``: physical ``[O][kh*kw][I]``), carved out of one symmetric-heap region (``wshadow``); the
  fp32 master copy and the optimizer state exist only on the parameter-server owner of each tile.  1-D
  parameters (BN, biases) stay fp32 (``vparams`` / ``vgrads`` regions).  Weight gradients are read by the
  coding kernels *where autograd leaves them* (bf16, same physical layout, address taken from a pointer
  table): no cast / re-layout / accumulate kernels between cuDNN and the coder.
* **Coding units.**  The reference matricizes a conv gradient ``(O,I,kh,kw)`` to ``(O*I/2, 2*kh*kw)``
  (``/root/reference/src/codings/svd.py:12-28``).  In the ``[O][K][I]`` layout one output channel is a
  contiguous *slab* of ``K*I`` values holding ``I/2`` rows of that matrix (row ``(o, ri)``, column ``(b, k)``
  = ``X_o[k][2*ri + b]``): kind ``SLAB``.  2-D tensors (fc, 1x1 convs) are handled in their tall orientation
  and cut into column blocks of <= 64 columns, each block an independent unit (kind ``MAT``) whose complete
  Gram/Jacobi SVD is exact — a block-spectral atomic decomposition (atoms of different blocks are orthogonal
  in the Frobenius inner product), so the ATOMO estimator stays exactly unbiased for square-ish layers
  without a truncated range finder.  Odd-``I`` convs (the 3-channel stem) and anything skinny travel dense.
* **Groups.**  Units are grouped by backward order; each group is encoded and pushed as soon as its last
  gradient exists, while backward continues (the reference's only overlap design,
  ``/root/reference/src/model_ops/resnet_split.py:259-360``).
* **Owners.**  Parameter-server work is sharded: PS tile ``j`` of a group belongs to owner ``j % n_owners``;
  a worker stores the ``U`` rows of a tile into that owner's arena only.  ``n_owners == 1`` is the
  reference's centralized PS.

Pure Python (unit-testable on CPU).  Struct layouts mirror ``csrc/v2_common.cuh``.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

KIND_SLAB, KIND_MAT, KIND_DENSE16, KIND_VEC = 1, 2, 3, 4
RCAP_MAX = 32
MAX_COLS = 64
BLOCK_COLS = 32               # column-block width of MAT units (Jacobi cost ~ cols^3 sits in the encode launch)
PS_TILE_ELEMS = 4608          # >= one slab of a 512-channel 3x3 conv
PS_MAX_ROWS = 256
DENSE_TILE_ELEMS = 4096
ENC_TILE_BYTES = 18432        # target bytes of gradient per encode tile
MAX_WORKERS = 16
MAX_GROUPS = 8
W_ALIGN = 64                  # bf16 elements (128 B)
V_ALIGN = 32                  # fp32 elements (128 B)

UNIT_FMT = "<4q20i"           # 112 bytes, mirrors struct Unit2
TILE_FMT = "<4i"              # unit, a, b, owner
UNIT_BYTES = struct.calcsize(UNIT_FMT)
TILE_BYTES = struct.calcsize(TILE_FMT)
CTRL2_FMT = "<iiffffiiQffffiiii"   # mirrors struct Ctrl2 (72 bytes)
CTRL2_BYTES = struct.calcsize(CTRL2_FMT)


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def slot_u_off(rcap: int, cols: int) -> int:
    return _round_up(4 + rcap + rcap * cols, 4)


def slot_floats(rows: int, cols: int, rcap: int, ubits: int = 0) -> int:
    """Slot size in floats.  ``ubits == 8`` (QSVD): U is stored as int8 [rows][rcap] followed by one fp32 scale per
    row instead of fp32 [rows][rcap]."""
    if ubits == 8:
        return _round_up(slot_u_off(rcap, cols) + _round_up(rows * rcap // 4, 4) + rows, 32)
    return _round_up(slot_u_off(rcap, cols) + rows * rcap, 32)


def slot_scale_off(rows: int, cols: int, rcap: int) -> int:
    """Float offset (inside the slot) of the per-row scales of an int8 U."""
    return slot_u_off(rcap, cols) + _round_up(rows * rcap // 4, 4)


def slot_capacity(cols: int, rank: int, systematic: bool) -> int:
    if rank <= 0:
        cap = cols
    elif systematic:
        cap = min(cols, rank)
    else:
        cap = min(cols, 2 * rank + 2)
    return min(_round_up(max(cap, 1), 4), RCAP_MAX)


@dataclass
class Param2:
    """One model parameter as the engine stores it."""
    index: int
    shape: Tuple[int, ...]
    is_w: bool              # bf16 shadow (dim >= 2) vs fp32 vector
    off: int                # element offset in wshadow (bf16) or vparams (fp32)
    numel: int
    widx: int = -1          # index among W params (gradient pointer table)
    group: int = 0

    def phys_strides(self) -> Tuple[int, ...]:
        """Element strides of the bf16 leaf tensor inside wshadow (channels_last for 4-D)."""
        s = self.shape
        if len(s) == 4:
            o, i, kh, kw = s
            return (i * kh * kw, 1, kw * i, i)
        st, acc = [], 1
        for d in reversed(s):
            st.append(acc)
            acc *= d
        return tuple(reversed(st))


@dataclass
class Unit2:
    index: int
    kind: int
    param: int              # Param2.index
    pidx: int               # gradient pointer table index (W params) or -1
    w_off: int              # element offset of the unit's base in wshadow / master (or vparams for VEC)
    g_off: int              # element offset inside the parameter's own gradient tensor
    rows: int = 0
    cols: int = 0
    K: int = 0
    I: int = 0
    rs: int = 0
    cs: int = 0
    rcap: int = 0
    budget: float = 0.0
    numel: int = 0
    group: int = 0
    slot_off: int = 0
    gpart_off: int = 0
    enc_tile0: int = 0
    n_enc: int = 0
    ps_rows: int = 0
    own0: int = 0
    ps_tile0: int = 0
    n_ps: int = 0
    ts_index: int = -1      # index among coded units (vsel / selcount / counters)
    ubits: int = 0          # 0: U stored fp32; 8: QSVD, U stochastically rounded to int8 with a per-row scale

    def pack(self) -> bytes:
        return struct.pack(UNIT_FMT, self.w_off, self.g_off, self.slot_off, self.gpart_off, self.kind, self.pidx,
                           self.rows, self.cols, self.K, self.I, self.rs, self.cs, self.rcap,
                           struct.unpack("<i", struct.pack("<f", self.budget))[0], self.numel, self.group,
                           self.enc_tile0, self.n_enc, self.ps_rows, self.own0, self.ps_tile0, self.n_ps,
                           self.ts_index, self.ubits)

    @property
    def coded(self) -> bool:
        return self.kind in (KIND_SLAB, KIND_MAT)


@dataclass
class Plan2:
    params: List[Param2]
    units: List[Unit2]
    enc_tiles: List[Tuple[int, int, int, int]]        # (unit, a, b, idx in unit)  SLAB: slab0, nslabs; MAT: row0, nrows;
    #                                                                       DENSE16: elem0, nelem (staging copy)
    ps_tiles: List[Tuple[int, int, int, int]]         # (unit, a, b, owner) sorted by (group, owner)
    enc_range: List[Tuple[int, int]]                  # per group: (first tile, count)
    ps_range: List[List[Tuple[int, int]]]             # [group][owner] -> (first tile, count)
    group_units: List[List[int]]
    n_groups: int
    n_owners: int
    w_total: int            # bf16 elements of wshadow
    v_total: int            # fp32 elements of vparams / vgrads
    stage_total: int        # bf16 elements of the dense-16 staging region
    arena_floats: int
    gpart_floats: int
    n_coded: int
    rank: int
    code: str

    def units_bytes(self) -> bytes:
        return b"".join(u.pack() for u in self.units)

    @staticmethod
    def tiles_bytes(tiles) -> bytes:
        return b"".join(struct.pack(TILE_FMT, *t) for t in tiles) or b"\0" * TILE_BYTES

    def factor_bytes_per_worker(self) -> int:
        return sum(4 * (4 + u.rcap + u.rcap * u.cols) + (u.rows * (u.rcap + 4) if u.ubits == 8 else 4 * u.rows * u.rcap)
                   for u in self.units if u.coded)

    def expected_factor_bytes(self) -> float:
        """Bytes actually stored per worker and step for the expected number of atoms (U is written in groups
        of 4 atoms)."""
        tot = 0.0
        for u in self.units:
            if u.coded:
                atoms = min(u.budget if u.budget > 0 else u.cols, u.cols)
                a4 = min(u.rcap, _round_up(int(atoms + 0.999), 4))
                tot += 4 * (4 + u.rcap + u.rcap * u.cols) + (u.rows * (a4 + 4) if u.ubits == 8 else 4 * u.rows * a4)
        return tot

    def dense_bytes(self) -> int:
        return sum((2 if u.kind == KIND_DENSE16 else 4) * u.numel for u in self.units if not u.coded)


def default_groups(shapes: Sequence[Sequence[int]], n_groups: int) -> List[int]:
    """Assign parameters to backward groups.  ``parameters()`` order ~ forward order, so the LAST parameters
    form group 0 (their gradients exist first).

    What matters is what is left on the critical path after the last ``wgrad``: the encode -> project -> PS chain
    of the FINAL group.  So the final group is only the first weight tensor of the network (for the CNNs here the
    3-channel stem, which travels dense: a staging copy + two PS tiles), the group before it is the next few
    percent of the weights (its chain hides behind the stem's backward), and the early groups — where almost all
    bytes are — close once the share still to come drops below ``0.5 * 0.3**g``.  A 1-D parameter (BN, bias)
    joins the group of the weight tensor that precedes it in forward order (its gradient exists earlier)."""
    numels = []
    for s in shapes:
        n = 1
        for d in s:
            n *= int(d)
        numels.append(n if len(s) >= 2 else 0)
    total = sum(numels) or 1
    n_groups = max(1, min(n_groups, MAX_GROUPS))
    w_idx = [i for i, n in enumerate(numels) if n > 0]
    groups = [0] * len(shapes)
    reserve_last = n_groups >= 3 and len(w_idx) >= n_groups
    body = n_groups - 1 if reserve_last else n_groups
    acc, g = 0, 0
    for i in range(len(shapes) - 1, -1, -1):
        groups[i] = g
        acc += numels[i]
        if numels[i] > 0 and g < body - 1 and (total - acc) / total <= 0.5 * 0.3 ** g:
            g += 1
    if reserve_last:
        groups[w_idx[0]] = g + 1
    last_w = None
    for i, s in enumerate(shapes):
        if len(s) >= 2:
            last_w = groups[i]
        elif last_w is not None:
            groups[i] = last_w
    used = sorted(set(groups))
    remap = {g: k for k, g in enumerate(used)}
    return [remap[g] for g in groups]


def build_plan2(shapes: Sequence[Sequence[int]], code: str = "svd", rank: int = 3, systematic: bool = False,
                n_owners: int = 1, n_groups: int = 4, groups: Optional[Sequence[int]] = None,
                block_cols: int = BLOCK_COLS, min_coded_numel: int = 256) -> Plan2:
    shapes = [tuple(int(d) for d in s) for s in shapes]
    ubits = 0
    if code == "qsvd":          # QSVD: spectral atoms with quantized left factors (README.md:141-142 of the reference)
        code, ubits = "svd", 8
    if groups is None:
        groups = default_groups(shapes, n_groups)
    n_groups = max(groups) + 1 if groups else 1
    params: List[Param2] = []
    w_off = v_off = 0
    widx = 0
    for i, s in enumerate(shapes):
        numel = 1
        for d in s:
            numel *= d
        if len(s) >= 2:
            params.append(Param2(i, s, True, w_off, numel, widx, groups[i]))
            w_off += _round_up(numel, W_ALIGN)
            widx += 1
        else:
            params.append(Param2(i, s, False, v_off, numel, -1, groups[i]))
            v_off += _round_up(numel, V_ALIGN)

    units: List[Unit2] = []
    stage_off = 0

    def add(u: Unit2):
        u.index = len(units)
        units.append(u)

    for p in params:
        if not p.is_w:
            add(Unit2(0, KIND_VEC, p.index, -1, p.off, 0, numel=p.numel, group=p.group))
            continue
        s = p.shape
        coded = code == "svd" and p.numel >= min_coded_numel
        if coded and len(s) == 4 and s[2] * s[3] > 1:
            o, i, kh, kw = s
            k = kh * kw
            if i % 16 == 0 and 2 * k <= MAX_COLS and o * i // 2 >= 2 * k and k * i <= PS_TILE_ELEMS:
                add(Unit2(0, KIND_SLAB, p.index, p.widx, p.off, 0, rows=o * i // 2, cols=2 * k, K=k, I=i,
                          rcap=slot_capacity(2 * k, rank, systematic), budget=float(rank), numel=p.numel,
                          group=p.group))
                continue
            coded = False
        if coded:
            # 2-D physical matrix [O][I] (Linear, 1x1 conv in channels_last): tall orientation, column blocks
            o = s[0]
            i = p.numel // o
            if o >= i:
                rows, cols, rs, cs = o, i, i, 1
            else:
                rows, cols, rs, cs = i, o, 1, i
            # a trailing 1-column block (cols % block_cols == 1) has no eigenproblem to solve: such tensors (none in
            # the model zoo) travel dense instead of exercising a degenerate unit in the kernels
            if cols >= 2 and cols % block_cols != 1:
                nb = (cols + block_cols - 1) // block_cols
                bud = float(rank) if nb == 1 else float(max(1, -(-rank // nb)))
                for b in range(nb):
                    c0 = b * block_cols
                    bc = min(block_cols, cols - c0)
                    add(Unit2(0, KIND_MAT, p.index, p.widx, p.off + c0 * cs, c0 * cs, rows=rows, cols=bc, rs=rs,
                              cs=cs, rcap=slot_capacity(bc, int(bud), systematic), budget=bud, numel=rows * bc,
                              group=p.group))
                continue
        u = Unit2(0, KIND_DENSE16, p.index, p.widx, p.off, 0, numel=p.numel, group=p.group)
        u.rs = stage_off                      # offset of the staging copy (bf16 elements)
        stage_off += _round_up(p.numel, W_ALIGN)
        add(u)

    # ---- tiles, slots ---------------------------------------------------------------------------------
    enc_tiles: List[Tuple[int, int, int, int]] = []
    ps_by_group: List[List[Tuple[int, int, int]]] = [[] for _ in range(n_groups)]
    enc_range, group_units = [], [[] for _ in range(n_groups)]
    slot_off = gpart_off = 0
    n_coded = 0
    for g in range(n_groups):
        first = len(enc_tiles)
        for u in units:
            if u.group != g:
                continue
            group_units[g].append(u.index)
            u.enc_tile0 = len(enc_tiles)
            if u.kind == KIND_SLAB:
                slab_bytes = u.K * u.I * 2
                ns = max(1, ENC_TILE_BYTES // slab_bytes)
                nslabs = u.rows // (u.I // 2)
                for s0 in range(0, nslabs, ns):
                    enc_tiles.append((u.index, s0, min(ns, nslabs - s0), s0 // ns))
                spt = max(1, min(PS_TILE_ELEMS // (u.K * u.I), PS_MAX_ROWS // (u.I // 2)))
                u.ps_rows = spt * (u.I // 2)
            elif u.kind == KIND_MAT:
                npad = (u.cols + 3) // 4 * 4
                er = max(64, min(1024, 8192 // npad) // 64 * 64)      # fp32 staging tile <= 32 KB
                for r0 in range(0, u.rows, er):
                    enc_tiles.append((u.index, r0, min(er, u.rows - r0), r0 // er))
                nc4 = (u.cols + 3) // 4
                u.ps_rows = max(8, min(PS_MAX_ROWS, PS_TILE_ELEMS // u.cols) & ~7)
            elif u.kind == KIND_DENSE16:
                for e0 in range(0, u.numel, 8192):     # staging copy tiles
                    enc_tiles.append((u.index, e0, min(8192, u.numel - e0), 0))
            u.n_enc = len(enc_tiles) - u.enc_tile0
            if u.coded:
                u.ts_index = n_coded
                n_coded += 1
                u.ubits = ubits
                u.slot_off = slot_off
                slot_off += slot_floats(u.rows, u.cols, u.rcap, ubits)
                u.gpart_off = gpart_off
                gpart_off += u.n_enc * u.cols * u.cols
                u.ps_tile0 = len(ps_by_group[g])
                for r0 in range(0, u.rows, u.ps_rows):
                    ps_by_group[g].append((u.index, r0, min(u.ps_rows, u.rows - r0)))
                u.n_ps = len(ps_by_group[g]) - u.ps_tile0
                u.own0 = u.ps_tile0 % n_owners
            else:
                u.ps_tile0 = len(ps_by_group[g])
                for e0 in range(0, u.numel, DENSE_TILE_ELEMS):
                    ps_by_group[g].append((u.index, e0, min(DENSE_TILE_ELEMS, u.numel - e0)))
                u.n_ps = len(ps_by_group[g]) - u.ps_tile0
                u.own0 = u.ps_tile0 % n_owners
        enc_range.append((first, len(enc_tiles) - first))

    ps_tiles: List[Tuple[int, int, int, int]] = []
    ps_range: List[List[Tuple[int, int]]] = []
    for g in range(n_groups):
        row = []
        for o in range(n_owners):
            first = len(ps_tiles)
            for j, (ui, a, b) in enumerate(ps_by_group[g]):
                if j % n_owners == o:
                    ps_tiles.append((ui, a, b, o))
            row.append((first, len(ps_tiles) - first))
        ps_range.append(row)
    return Plan2(params, units, enc_tiles, ps_tiles, enc_range, ps_range, group_units, n_groups, n_owners,
                 max(w_off, W_ALIGN), max(v_off, V_ALIGN), max(stage_off, W_ALIGN), max(slot_off, 32),
                 max(gpart_off, 1), n_coded, rank, code)


def owner_of_row(u: Unit2, row: int, n_owners: int) -> int:
    """Owner of the PS tile that holds tall row ``row`` of a coded unit (what project_push computes)."""
    return (u.own0 + row // u.ps_rows) % n_owners


OPT_SGD, OPT_ADAM, OPT_AMSGRAD = 0, 1, 2


def pack_ctrl2(step: int = 1, lr: float = 0.01, momentum: float = 0.0, dampening: float = 0.0,
               weight_decay: float = 0.0, nesterov: bool = False, first_step: int = 1, seed: int = 1,
               beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, opt: int = OPT_SGD,
               num_aggregate: int = 0) -> bytes:
    return struct.pack(CTRL2_FMT, step, 0, lr, momentum, dampening, weight_decay, int(nesterov), first_step,
                       seed & 0xFFFFFFFFFFFFFFFF, beta1, beta2, eps, 0.0, int(opt), int(num_aggregate), 0, 0)
