"""Host-side mirror of the reference's place-recognition interface for the retrieval
path, backed by libsemgate's sm_100a kernels.

Same public names, argument meaning and error behaviour as the reference's
scripts/semantic_gating/place_recognition.py (PlaceMatch :61-69, PlaceDescriptor
:72-78, BasePlaceRecognition :81-190, SemanticPlaceRecognition :806-933), so code
written against the reference keeps working:

    spr = SemanticPlaceRecognition(vpr_method='mixvpr', device='cuda')
    spr.add_image(descriptor, timestamp, floor_label)     # or append PlaceDescriptor
    matches = spr.find_loop_closures(enable_floor_gating=True, k=10)

What differs by design
  * Descriptor *extraction* (MixVPR / SALAD / AnyLoc / CricaVPR model inference,
    :193-803) is out of scope: `extract_descriptor` accepts an already extracted
    descriptor vector of the method's dimensionality and returns it.
  * The database lives on the GPU as row-normalised bf16 (packed incrementally, not
    re-stacked per call as :140,:177 do) and the N x N matrix is never formed.
  * `find_loop_closures(..., gate_mode="mask")` is an extension: cross-floor columns
    are excluded before top-k.  The default "flag" keeps the reference's order
    (top-k, threshold, then flag; :888-899).
  * Tie order among exactly equal scores is (lower index first); the reference's
    `np.argsort` leaves it unspecified.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native

GATE_MODES = {"flag": _native.GATE_FLAG, "mask": _native.GATE_MASK}


@dataclass
class PlaceMatch:
    """A place-recognition match between two keyframes."""
    query_idx: int
    match_idx: int
    similarity: float
    query_timestamp: Optional[float] = None
    match_timestamp: Optional[float] = None
    is_valid: bool = True  # cleared by the floor gate


@dataclass
class PlaceDescriptor:
    """Global descriptor of one keyframe."""
    timestamp: float
    descriptor: np.ndarray
    image_path: Optional[str] = None
    floor_label: Optional[int] = None


@dataclass
class MatchArrays:
    """Struct-of-arrays form of a match list (query ascending, similarity descending)."""
    query_idx: np.ndarray
    match_idx: np.ndarray
    similarity: np.ndarray
    is_valid: np.ndarray
    query_timestamp: Optional[np.ndarray] = None
    match_timestamp: Optional[np.ndarray] = None

    def __len__(self):
        return int(self.query_idx.shape[0])

    def to_matches(self) -> List[PlaceMatch]:
        qt = self.query_timestamp.tolist() if self.query_timestamp is not None else [None] * len(self)
        mt = self.match_timestamp.tolist() if self.match_timestamp is not None else [None] * len(self)
        return [PlaceMatch(q, m, s, a, b, v) for q, m, s, a, b, v in
                zip(self.query_idx.tolist(), self.match_idx.tolist(), self.similarity.astype(np.float64).tolist(),
                    qt, mt, self.is_valid.astype(bool).tolist())]

    def as_gate_candidates(self) -> List[Tuple[int, int, float]]:
        """`(query_idx, match_idx, score)` tuples, the input of SemanticLoopClosureGate.gate_candidates."""
        return list(zip(self.query_idx.tolist(), self.match_idx.tolist(), self.similarity.astype(np.float64).tolist()))


def _effective_k(k, n: int) -> int:
    """The reference slices `argsort()[::-1][:k]` (place_recognition.py:150,888): any k >= 0 works and more than
    n candidates never come back.  Here k <= 64 is one sweep, larger k several (up to MAX_K_TOTAL)."""
    k = int(k)
    if k < 0:
        raise ValueError(f"k={k} must not be negative")
    k = min(k, int(n))
    if k > _native.MAX_K_TOTAL:
        raise ValueError(f"k={k} exceeds the library's limit {_native.MAX_K_TOTAL}")
    return k


def _device_index(device) -> int:
    if isinstance(device, int):
        return device
    s = str(device)
    if s == "cpu":
        raise RuntimeError("semgate has no CPU path: the retrieval kernels are sm_100a only (use device='cuda').")
    if ":" in s:
        return int(s.split(":")[1])
    try:
        import torch
        return torch.cuda.current_device() if torch.cuda.is_available() else 0
    except Exception:
        return 0


class _PackedDB:
    """Device-resident database: bf16 normalised rows, fp64 timestamps, int32 floors, grown geometrically and
    re-synchronised from the Python descriptor list on every call (callers may append, replace or remove
    PlaceDescriptor objects directly, place_recognition.py:1015-1020; the reference re-reads everything per call,
    :140, :177, :870-899).

    What is checked per call, for EVERY element (O(n) of host work, no device work unless something changed):
    the identity of each PlaceDescriptor and of its `descriptor` array (strong references to both are held, so
    an address cannot be recycled behind our back) -> rows from the first difference on are re-normalised and
    re-uploaded; every `timestamp` and `floor_label` value -> re-uploaded when any differs (12 B per keyframe).
    What cannot be seen: the CONTENTS of a descriptor array edited in place.  Call `invalidate()` after that."""

    def __init__(self, engine: "_native.Engine"):
        self.engine = engine
        self.n = 0
        self.d = None
        self.cap = 0
        self.bf16 = self.ts = self.floor = None
        self._refs: List[PlaceDescriptor] = []      # the records the device rows were built from
        self._drefs: List[np.ndarray] = []          # ... and their descriptor arrays
        self._ts_host = np.zeros(0, np.float64)
        self._fl_host = np.zeros(0, np.int32)
        self.has_floor = False

    def invalidate(self):
        """Forget the device copy: the next call re-reads and re-uploads every descriptor."""
        self.n = 0
        self._refs, self._drefs = [], []
        self._ts_host, self._fl_host = np.zeros(0, np.float64), np.zeros(0, np.int32)

    def _grow(self, need: int, d: int):
        import torch
        if self.d is not None and self.d != d:
            if self.n:
                raise ValueError(f"descriptor length changed from {self.d} to {d}")
            self.cap = 0                       # nothing packed: a new length may start over
        if need <= self.cap:
            self.d = d
            return
        cap = max(need, int(self.cap * 1.5), 1024)
        dev = torch.device("cuda", self.engine.device)
        dp = _native.pad_dim(d)
        bf16 = torch.empty((cap, dp), dtype=torch.bfloat16, device=dev)
        ts = torch.empty((cap,), dtype=torch.float64, device=dev)
        fl = torch.empty((cap,), dtype=torch.int32, device=dev)
        if self.n:
            bf16[:self.n].copy_(self.bf16[:self.n])
        self.bf16, self.ts, self.floor, self.cap, self.d = bf16, ts, fl, cap, d
        self._ts_host, self._fl_host = np.zeros(0, np.float64), np.zeros(0, np.int32)   # new buffers: upload again

    @staticmethod
    def _first_difference(old: list, new: list) -> int:
        """Length of the common prefix of two object lists, by identity."""
        m = min(len(old), len(new))
        if m == 0:
            return 0
        a = np.fromiter(map(id, old[:m]), dtype=np.int64, count=m)
        b = np.fromiter(map(id, new[:m]), dtype=np.int64, count=m)
        diff = np.nonzero(a != b)[0]
        return int(diff[0]) if diff.size else m

    def sync(self, descriptors: List[PlaceDescriptor]):
        import torch
        n = len(descriptors)
        drefs = [p.descriptor for p in descriptors]
        keep = min(self.n, self._first_difference(self._refs, descriptors), self._first_difference(self._drefs, drefs))
        if n > keep:
            new = drefs[keep:]
            x = np.vstack([np.asarray(a).reshape(1, -1) for a in new])
            # fp16 descriptors (extractors under autocast) cross PCIe as they are: K1 widens them exactly
            x = np.ascontiguousarray(x, dtype=np.float16 if x.dtype == np.float16 else np.float32)
            d = x.shape[1]
            self.n = keep                      # rows beyond `keep` are stale: _grow must not carry them over
            self._grow(n, d)
            xt = torch.from_numpy(x).to(self.bf16.device, non_blocking=False)
            self.engine.normalize_cast(xt, out=self.bf16[keep:n])
        self.n = n
        self._refs, self._drefs = list(descriptors), drefs
        if n == 0:
            return
        # timestamps and labels: read every one, upload when anything differs
        ts = np.fromiter((float(p.timestamp) for p in descriptors), dtype=np.float64, count=n)
        fl64 = np.fromiter((_native.FLOOR_NONE if p.floor_label is None else int(p.floor_label) for p in descriptors),
                           dtype=np.int64, count=n)
        real = fl64[fl64 != _native.FLOOR_NONE]
        if real.size and (real.max() > 2**31 - 1 or real.min() < -2**31 + 1):
            raise ValueError("floor labels must fit in int32")
        fl = fl64.astype(np.int32)
        dev = self.bf16.device
        if not np.array_equal(ts, self._ts_host, equal_nan=True):
            self.ts[:n] = torch.from_numpy(ts).to(dev)
            self._ts_host = ts
        if not np.array_equal(fl, self._fl_host):
            self.floor[:n] = torch.from_numpy(fl).to(dev)
            self._fl_host = fl


class BasePlaceRecognition:
    """Base class of the VPR methods (reference :81-190)."""

    def __init__(self, descriptor_dim: int = 4096, device: str = 'cuda'):
        self.descriptor_dim = descriptor_dim
        self.device = device
        self.model = None
        self.descriptors: List[PlaceDescriptor] = []
        self._db: Optional[_PackedDB] = None

    # -- engine plumbing ---------------------------------------------------
    def _engine(self) -> "_native.Engine":
        return _native.get_engine(_device_index(self.device))

    def _packed(self) -> _PackedDB:
        if self._db is None:
            self._db = _PackedDB(self._engine())
        self._db.sync(self.descriptors)
        return self._db

    def invalidate(self):
        """Drop the device copy of the database.  Needed only after editing the CONTENTS of a descriptor array
        in place: appended, replaced or removed PlaceDescriptor objects, replaced `descriptor` arrays and changed
        `timestamp` / `floor_label` values are picked up by every call on their own."""
        if self._db is not None:
            self._db.invalidate()

    # -- reference interface ---------------------------------------------------
    def extract_descriptor(self, image: np.ndarray) -> np.ndarray:
        """Model inference is outside this package: pass the extracted descriptor."""
        a = np.asarray(image)
        if a.ndim == 1 and (self.descriptor_dim is None or a.shape[0] == self.descriptor_dim):
            return a
        raise NotImplementedError(
            "descriptor extraction is not part of semgate; pass a 1-D descriptor of length "
            f"{self.descriptor_dim} (got shape {a.shape})")

    def add_image(self, image: np.ndarray, timestamp: float, floor_label: Optional[int] = None,
                  image_path: Optional[str] = None) -> PlaceDescriptor:
        descriptor = self.extract_descriptor(image)
        pd = PlaceDescriptor(timestamp=timestamp, descriptor=descriptor, image_path=image_path, floor_label=floor_label)
        self.descriptors.append(pd)
        return pd

    def query(self, image: np.ndarray, timestamp: Optional[float] = None, k: int = 5,
              min_time_gap: float = 10.0) -> List[PlaceMatch]:
        """Top-k most similar database keyframes (reference :117-163): temporal mask only
        when `timestamp` is given, masked entries dropped, `query_idx = len(descriptors)`."""
        if len(self.descriptors) == 0:
            return []
        q = np.asarray(self.extract_descriptor(image), dtype=np.float32).reshape(1, -1)
        res = self.query_batch(q, None if timestamp is None else np.array([timestamp], dtype=np.float64), k, min_time_gap)
        scores, idx, count = res
        c = int(count[0])
        nq = len(self.descriptors)
        return [PlaceMatch(query_idx=nq, match_idx=int(idx[0, i]), similarity=float(scores[0, i]),
                           query_timestamp=timestamp, match_timestamp=self.descriptors[int(idx[0, i])].timestamp)
                for i in range(c)]

    def query_batch(self, queries: np.ndarray, timestamps: Optional[np.ndarray] = None, k: int = 5,
                    min_time_gap: float = 10.0):
        """Batched `query`: (scores[nq,k], idx[nq,k], count[nq]) numpy arrays."""
        import torch
        db = self._packed()
        eng = db.engine
        dev = db.bf16.device
        q = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)).to(dev)
        if q.shape[1] != db.d:
            raise ValueError(f"query length {q.shape[1]} != database descriptor length {db.d}")
        qb = eng.normalize_cast(q)
        k = _effective_k(k, db.n)
        if k == 0:
            nq = q.shape[0]
            return np.zeros((nq, 0), np.float32), np.zeros((nq, 0), np.int32), np.zeros((nq,), np.int32)
        params = _native.make_params(k=k, similarity_threshold=-np.inf, min_time_gap=min_time_gap, max_floor_diff=-1)
        q_ts = db_ts = None
        if timestamps is not None:
            q_ts = torch.from_numpy(np.ascontiguousarray(timestamps, dtype=np.float64)).to(dev)
            db_ts = db.ts[:db.n]
        r = eng.gated_topk(qb, db.bf16[:db.n], params, q_ts=q_ts, db_ts=db_ts)
        return r.scores.cpu().numpy(), r.idx.cpu().numpy(), r.count.cpu().numpy()

    def _compute_similarity(self, query: np.ndarray, database: np.ndarray) -> np.ndarray:
        """Cosine similarity of one query against every row of `database` (reference :165-171):
        both sides normalised as `x / (|x| + 1e-8)`, fp32 result of length N."""
        import torch
        database = np.ascontiguousarray(database, dtype=np.float32)
        if database.ndim != 2 or database.shape[0] == 0:
            return np.zeros((0,), dtype=np.float32)
        eng = self._engine()
        dev = torch.device("cuda", eng.device)
        q = np.ascontiguousarray(np.asarray(query, dtype=np.float32).reshape(1, -1))
        if q.shape[1] != database.shape[1]:
            raise ValueError("shapes not aligned")           # numpy's dot raises ValueError too
        qb = eng.normalize_cast(torch.from_numpy(q).to(dev))
        db = eng.normalize_cast(torch.from_numpy(database).to(dev))
        return eng.similarity_matrix(qb, db)[0].cpu().numpy()

    def build_descriptor_matrix(self) -> np.ndarray:
        if len(self.descriptors) == 0:
            return np.array([])
        return np.vstack([d.descriptor for d in self.descriptors])

    def compute_all_pairwise_similarities(self) -> np.ndarray:
        """N x N cosine similarities (reference :179-190) from the packed device database.  The
        retrieval path (find_loop_closures / query) never forms this matrix; here it is the product,
        so it must fit in device and host memory (4*N*N bytes)."""
        import torch
        n = len(self.descriptors)
        if n == 0:
            return np.array([])
        db = self._packed()
        free, _ = torch.cuda.mem_get_info(db.bf16.device)
        if 4 * n * n > free:
            raise MemoryError(f"a {n} x {n} fp32 similarity matrix does not fit in device memory; "
                              "use find_loop_closures / query, which never form it")
        return db.engine.similarity_matrix(db.bf16[:n], db.bf16[:n]).cpu().numpy()


class MixVPR(BasePlaceRecognition):
    def __init__(self, device: str = 'cuda', **_):
        super().__init__(descriptor_dim=4096, device=device)   # reference :197


class SALAD(BasePlaceRecognition):
    def __init__(self, device: str = 'cuda', **_):
        super().__init__(descriptor_dim=8448, device=device)   # reference :340


class AnyLoc(BasePlaceRecognition):
    def __init__(self, device: str = 'cuda', **_):
        super().__init__(descriptor_dim=49152, device=device)  # reference :418


class _LocalStore:
    """Device-resident patch features: bf16 [cap, P, pad64(D)], rows L2-normalised once per
    keyframe (place_recognition.py:695-699 re-normalises them for every pair)."""

    def __init__(self, engine: "_native.Engine"):
        self.engine = engine
        self.slot: Dict[int, int] = {}     # keyframe index -> row of `feats`
        self._ids: Dict[int, np.ndarray] = {}   # keyframe index -> the cached array itself (a held reference: an
                                                # address cannot be recycled for a replacement array)
        self.feats = None
        self.P = self.D = None

    def sync(self, cache: Dict[int, np.ndarray]):
        import torch
        new = [k for k, v in cache.items() if self._ids.get(k) is not v]
        if not new:
            return
        first = np.asarray(cache[new[0]])
        P, D = first.shape[-2], first.shape[-1]
        if self.P is not None and (P, D) != (self.P, self.D):
            raise ValueError(f"local feature shape changed from {(self.P, self.D)} to {(P, D)}")
        self.P, self.D = P, D
        need = len(self.slot) + sum(1 for k in new if k not in self.slot)
        dev = torch.device("cuda", self.engine.device)
        dp = _native.pad_dim(D)
        if self.feats is None or self.feats.shape[0] < need:
            cap = max(need, 64, 0 if self.feats is None else int(self.feats.shape[0] * 1.5))
            grown = torch.empty((cap, P, dp), dtype=torch.bfloat16, device=dev)
            if self.feats is not None and len(self.slot):
                grown[:self.feats.shape[0]].copy_(self.feats)
            self.feats = grown
        for k in new:
            a = np.asarray(cache[k], dtype=np.float32).reshape(-1, D)
            if a.shape[0] != P:
                raise ValueError("every keyframe needs the same number of patches")
            if k not in self.slot:
                self.slot[k] = len(self.slot)
            self.engine.normalize_cast(torch.from_numpy(np.ascontiguousarray(a)).to(dev), out=self.feats[self.slot[k]])
            self._ids[k] = cache[k]


class CricaVPR(BasePlaceRecognition):
    """CricaVPR interface (reference :508-803) for the retrieval + cross-correlation re-rank part.
    Model inference is out of scope: global descriptors and patch-level local features are
    passed in (`add_image(descriptor, ..., local_features=patches)`)."""

    def __init__(self, device: str = 'cuda', use_reranking: bool = True, **_):
        super().__init__(descriptor_dim=10752, device=device)  # reference :513
        self.use_reranking = use_reranking
        self._feature_cache: Dict[int, np.ndarray] = {}
        self._store: Optional[_LocalStore] = None

    def extract_local_features(self, image: np.ndarray) -> np.ndarray:
        """Accepts already extracted patch features `[P, D]` / `[1, P, D]`; returns `[1, P, D]` (reference :655-667)."""
        a = np.asarray(image)
        if a.ndim == 2:
            return a[None]
        if a.ndim == 3 and a.shape[0] == 1:
            return a
        raise NotImplementedError("local feature extraction is not part of semgate; pass [P, D] patch features")

    def add_image(self, image: np.ndarray, timestamp: float, floor_label: Optional[int] = None,
                  image_path: Optional[str] = None, local_features: Optional[np.ndarray] = None) -> PlaceDescriptor:
        pd = super().add_image(image, timestamp, floor_label, image_path)
        if self.use_reranking and local_features is not None:
            self._feature_cache[len(self.descriptors) - 1] = self.extract_local_features(local_features)
        return pd

    def _local(self) -> _LocalStore:
        if self._store is None:
            self._store = _LocalStore(self._engine())
        self._store.sync(self._feature_cache)
        return self._store

    def _pair_scores(self, q_idx: np.ndarray, m_idx: np.ndarray, global_sim: np.ndarray):
        """(cross, combined) fp32 arrays for keyframe-index pairs; indices without cached features
        get cross = NaN and combined = global (reference :741-749)."""
        import torch
        st = self._local()
        dev = torch.device("cuda", st.engine.device)
        slot = st.slot
        qs = np.array([slot.get(int(i), -1) for i in q_idx], dtype=np.int32)
        ms = np.array([slot.get(int(i), -1) for i in m_idx], dtype=np.int32)
        if st.feats is None:
            return np.full(len(qs), np.nan, np.float32), np.asarray(global_sim, dtype=np.float32).copy()
        cross, comb = st.engine.rerank_scores(st.feats, torch.from_numpy(qs).to(dev), torch.from_numpy(ms).to(dev),
                                              torch.from_numpy(np.ascontiguousarray(global_sim, dtype=np.float32)).to(dev))
        return cross.cpu().numpy(), comb.cpu().numpy()

    def compute_cross_correlation_score(self, query_features: np.ndarray, match_features: np.ndarray) -> float:
        """Cross-correlation score of two patch-feature sets (reference :669-710)."""
        import torch
        eng = self._engine()
        dev = torch.device("cuda", eng.device)
        q = np.asarray(query_features, dtype=np.float32)
        m = np.asarray(match_features, dtype=np.float32)
        q = q.reshape(-1, q.shape[-1])
        m = m.reshape(-1, m.shape[-1])
        if q.shape != m.shape:
            raise ValueError("the kernel scores equally shaped patch sets (same P and D)")
        both = torch.from_numpy(np.ascontiguousarray(np.concatenate([q, m]))).to(dev)
        feats = eng.normalize_cast(both).view(2, q.shape[0], -1)
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        cross, _ = eng.rerank_scores(feats, z, z + 1, torch.zeros(1, dtype=torch.float32, device=dev))
        return float(cross.item())

    def rerank_candidates(self, query_idx: int, candidates: List[Tuple[int, float]], top_k: int = 5) -> List[Tuple[int, float]]:
        """Re-rank `(match_idx, global_similarity)` candidates of one query (reference :712-757)."""
        if not self.use_reranking or query_idx not in self._feature_cache or not candidates:
            return candidates[:top_k]
        m = np.array([c[0] for c in candidates], dtype=np.int64)
        g = np.array([c[1] for c in candidates], dtype=np.float32)
        _, comb = self._pair_scores(np.full(len(m), query_idx), m, g)
        order = sorted(range(len(m)), key=lambda i: comb[i], reverse=True)      # stable, like list.sort (:754)
        return [(int(m[i]), float(comb[i])) for i in order[:top_k]]

    def rerank_batch(self, query_idx: np.ndarray, cand_idx: np.ndarray, global_sim: np.ndarray, count: np.ndarray,
                     top_k: int = 5):
        """All queries at once: padded `[Q, kc]` candidate lists (kc <= 64) -> (idx [Q,top_k], score [Q,top_k],
        count [Q]).  Queries without cached features keep their input order."""
        import torch
        st = self._local()
        eng = st.engine
        dev = torch.device("cuda", eng.device)
        Q, kc = cand_idx.shape
        live = np.arange(kc)[None, :] < np.asarray(count)[:, None]
        qq = np.repeat(np.asarray(query_idx, dtype=np.int64)[:, None], kc, axis=1)
        has_q = np.array([int(q) in st.slot for q in query_idx])
        m = np.where(live, cand_idx, -1)
        m = np.where(has_q[:, None] & self.use_reranking, m, -1)       # no query features: global score only
        _, comb = self._pair_scores(qq.ravel(), m.ravel(), np.where(live, global_sim, 0).astype(np.float32).ravel())
        comb = comb.reshape(Q, kc)
        # rows that must keep their order (reference :733-737): give them strictly decreasing keys
        keep = ~(has_q & self.use_reranking)
        keys = np.where(keep[:, None], -np.arange(kc, dtype=np.float32)[None, :], comb)
        oi, _, oc = eng.rerank_select(torch.from_numpy(np.ascontiguousarray(cand_idx, dtype=np.int32)).to(dev),
                                      torch.from_numpy(np.ascontiguousarray(keys, dtype=np.float32)).to(dev),
                                      torch.from_numpy(np.ascontiguousarray(count, dtype=np.int32)).to(dev), top_k)
        oi, oc = oi.cpu().numpy(), oc.cpu().numpy()
        # scores of the selected entries (for kept-order rows the global similarity)
        pos = {(r, int(c)): j for r in range(Q) for j, c in enumerate(cand_idx[r, :count[r]])}
        src = np.where(keep[:, None], global_sim, comb)
        os_ = np.full((Q, top_k), -np.inf, np.float32)
        for r in range(Q):
            for t in range(oc[r]):
                os_[r, t] = src[r, pos[(r, int(oi[r, t]))]]
        return oi, os_, oc


_METHODS = {'mixvpr': MixVPR, 'salad': SALAD, 'anyloc': AnyLoc, 'cricavpr': CricaVPR}


class SemanticPlaceRecognition:
    """Place recognition with floor gating (reference :806-933)."""

    def __init__(self, vpr_method: str = 'mixvpr', device: str = 'cuda', similarity_threshold: float = 0.5,
                 min_time_gap: float = 10.0, descriptor_dim: Optional[int] = None):
        self.similarity_threshold = similarity_threshold
        self.min_time_gap = min_time_gap
        cls = _METHODS.get(vpr_method.lower())
        if cls is None:
            raise ValueError(f"Unknown VPR method: {vpr_method}. Available: mixvpr, salad, anyloc, cricavpr")
        self.vpr = cls(device=device)
        if descriptor_dim is not None:
            self.vpr.descriptor_dim = descriptor_dim

    def add_image(self, image: np.ndarray, timestamp: float, floor_label: int,
                  image_path: Optional[str] = None) -> PlaceDescriptor:
        return self.vpr.add_image(image, timestamp, floor_label, image_path)

    def find_loop_closures(self, enable_floor_gating: bool = True, k: int = 10,
                           gate_mode: str = "flag") -> List[PlaceMatch]:
        """All loop-closure candidates of the database (reference :851-911)."""
        return self.find_loop_closures_arrays(enable_floor_gating, k, gate_mode).to_matches()

    def find_loop_closures_arrays(self, enable_floor_gating: bool = True, k: int = 10,
                                  gate_mode: str = "flag", valid_only: bool = False,
                                  with_statistics: bool = False):
        """Same result as struct-of-arrays (no per-match Python objects).
        `valid_only`: only the floor-consistent candidates (what geometric verification would keep,
        geometric_verification.py:709).  `with_statistics`: also return `get_statistics` of the full
        candidate list, reduced on the device (returns `(arrays, stats)`)."""
        e = np.zeros(0, dtype=np.int32)
        empty = MatchArrays(e, e.copy(), np.zeros(0, np.float32), np.zeros(0, bool), np.zeros(0), np.zeros(0))
        if len(self.vpr.descriptors) < 2:
            return (empty, self.get_statistics([])) if with_statistics else empty
        if gate_mode not in GATE_MODES:
            raise ValueError(f"gate_mode must be one of {sorted(GATE_MODES)}")
        db = self.vpr._packed()
        eng = db.engine
        n = db.n
        k = _effective_k(k, n)
        if k == 0:
            return (empty, self.get_statistics([])) if with_statistics else empty
        params = _native.make_params(k=int(k), similarity_threshold=self.similarity_threshold,
                                     min_time_gap=self.min_time_gap,
                                     max_floor_diff=0 if enable_floor_gating else -1, gate_mode=GATE_MODES[gate_mode])
        ts, fl = db.ts[:n], db.floor[:n]
        r = eng.gated_topk(db.bf16[:n], db.bf16[:n], params, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl)
        stats = None
        if with_statistics:
            aq, am, as_, av, atot = eng.compact(r)
            stats = self._stats_dict(eng.candidate_stats(as_, av, atot).cpu().numpy())
            if not valid_only:
                oq, om, os_, ov, total = aq, am, as_, av, atot
        if valid_only or not with_statistics:
            oq, om, os_, ov, total = eng.compact(r, valid_only=valid_only)
        t = int(total.item())
        q = oq[:t].cpu().numpy()
        m = om[:t].cpu().numpy()
        tsh = ts.cpu().numpy()
        arrays = MatchArrays(q, m, os_[:t].cpu().numpy(), ov[:t].cpu().numpy().astype(bool), tsh[q], tsh[m])
        return (arrays, stats) if with_statistics else arrays

    @staticmethod
    def _stats_dict(v) -> Dict:
        """(total, valid, sum, sum_valid) -> the reference's statistics dict (:913-933)."""
        n, valid = int(v[0]), int(v[1])
        if n == 0:
            return {'total_matches': 0, 'valid_matches': 0, 'rejected_matches': 0, 'rejection_rate': 0.0}
        return {'total_matches': n, 'valid_matches': valid, 'rejected_matches': n - valid, 'rejection_rate': (n - valid) / n,
                'mean_similarity': float(v[2]) / n, 'mean_valid_similarity': float(v[3]) / valid if valid > 0 else 0.0}

    def get_statistics(self, matches) -> Dict:
        """Statistics of a match list (reference :913-933).  Accepts List[PlaceMatch] or MatchArrays."""
        if isinstance(matches, MatchArrays):
            valid_mask = matches.is_valid.astype(bool)
            sims = matches.similarity.astype(np.float64)
        else:
            valid_mask = np.array([m.is_valid for m in matches], dtype=bool)
            sims = np.array([m.similarity for m in matches], dtype=np.float64)
        n = int(valid_mask.shape[0])
        if n == 0:
            return {'total_matches': 0, 'valid_matches': 0, 'rejected_matches': 0, 'rejection_rate': 0.0}
        valid = int(valid_mask.sum())
        return {
            'total_matches': n,
            'valid_matches': valid,
            'rejected_matches': n - valid,
            'rejection_rate': (n - valid) / n,
            'mean_similarity': np.mean(sims),
            'mean_valid_similarity': np.mean(sims[valid_mask]) if valid > 0 else 0.0,
        }
