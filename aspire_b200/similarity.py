"""Plugin boundary of the evaluation code: ``SimilarityModel`` and the otAspire / tsAspire models behind it.

Mirrors src/evaluation/utils/models.py:
  * ``SimilarityModel``  (:23-167)  -- same constructor, abstract ``encode`` / ``get_similarity``, cache helpers
  * ``AspireModel``      (:169-209) -- otAspire: ``encode`` = prepare_abstracts + AspireConSent.forward,
                                        ``get_similarity(x, y)`` = -OT_eps(x, y)
  * ``get_model``        (:738-768) -- factory for the model names on this path
and adds the batched entry point the reference lacks: ``score_pool(query_enc, cand_encs)`` scores one query
against a whole candidate pool in one kernel launch (the reference's evaluate.py:72-74 loop calls
``get_similarity`` once per pair; pp_gen_nearest.py:182-202 batches 64 at a time through caching_score).
"""
import logging
import os
from abc import ABCMeta, abstractmethod
from typing import Dict, List, Union

import numpy as np
import torch
from torch import Tensor

from . import _abi
from .consent import AspireConSent, prepare_abstracts, prepare_abstracts_native
from .distances import (AllPairMaskedWasserstein, bbox_diameter, epsilon_schedule, l2max_scores, ot_scores,
                        rep_len_tup)


_HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"


def _h5py_or_none():
    try:
        import h5py  # the reference's cache format (utils/models.py:76-81); optional here
        return h5py
    except ImportError:
        return None


class EncodingsCache:
    """{paper_id: fp32 encoding [S, D]} -- the encodings cache of src/evaluation/utils/models.py:68-124.

    The reference's cache is an ``h5py.File(cache_filename, 'a')`` with one dataset per paper id (:76-95, read back at
    :112-114).  When h5py is importable this class opens exactly that file (same open modes, same fallback to 'w' on
    a corrupt file), so caches written by the reference are read and vice versa.  Without h5py the cache persists as
    one ``.npz`` beside the requested name (``encodings.h5`` -> ``encodings.h5.npz``; the SAME path is used to load
    and to save), written only when something changed; an existing HDF5 file that cannot be read without h5py raises
    instead of being silently re-encoded.  Members: the ones the callers touch -- ``in``, ``get``, ``[]``, ``keys``,
    ``create_dataset(name=, data=)``, ``close``.
    """

    def __init__(self, filename=None):
        self.filename = filename
        self._mem, self._h5, self._dirty, self.path = {}, None, False, None
        if not filename:
            return
        h5py = _h5py_or_none()
        if h5py is not None and not filename.endswith(".npz"):
            try:
                self._h5 = h5py.File(filename, 'a')
            except Exception:
                logging.info(f"Error: could not open encodings cache {filename}.\nOverwriting the cache.")
                self._h5 = h5py.File(filename, 'w')
            return
        self.path = filename if filename.endswith(".npz") else filename + ".npz"
        if self.path != filename and os.path.exists(filename):
            with open(filename, "rb") as fh:
                if fh.read(8) == _HDF5_MAGIC:
                    raise _abi.AspireB200Error(f"{filename} is an HDF5 encodings cache; install h5py to read it")
        if os.path.exists(self.path):
            with np.load(self.path, allow_pickle=False) as z:
                names = [str(k) for k in z["__keys__"]] if "__keys__" in z.files else None
                if names is None:  # archives of the first layout: one entry per paper id
                    self._mem = {k: z[k] for k in z.files}
                else:
                    self._mem = {name: z[f"a{i}"] for i, name in enumerate(names)}

    def __contains__(self, pid):
        return pid in (self._h5 if self._h5 is not None else self._mem)

    def __len__(self):
        return len(self._h5 if self._h5 is not None else self._mem)

    def keys(self):
        return (self._h5 if self._h5 is not None else self._mem).keys()

    def get(self, pid, default=None):
        store = self._h5 if self._h5 is not None else self._mem
        return store[pid] if pid in store else default

    def __getitem__(self, pid):
        return (self._h5 if self._h5 is not None else self._mem)[pid]

    def create_dataset(self, name, data):
        arr = data.detach().cpu().numpy() if isinstance(data, Tensor) else np.asarray(data)
        if self._h5 is not None:
            self._h5.create_dataset(name=name, data=arr)
        else:
            self._mem[name] = arr
            self._dirty = True

    def close(self):
        if self._h5 is not None:
            self._h5.close()
            self._h5 = None
        elif self.path and self._dirty:
            names = list(self._mem)  # paper ids are data, not keyword names: stored as an array beside a0, a1, ...
            np.savez(self.path, __keys__=np.asarray(names, dtype=str), **{f"a{i}": self._mem[n] for i, n in enumerate(names)})
            self._dirty = False


def batchify(dataset: Dict, batch_size: int):
    """src/evaluation/utils/utils.py:10-27 -- (pids, papers) groups of ``batch_size`` in dict order."""
    items = list(dataset.items())
    for s in range(0, len(items), batch_size):
        chunk = items[s:s + batch_size]
        yield [p for p, _ in chunk], [d for _, d in chunk]


class SimilarityModel(metaclass=ABCMeta):
    """Abstract paper-similarity model: implement ``encode`` and ``get_similarity`` (higher == closer)."""
    ENCODING_TYPES = ('abstract', 'sentence', 'sentence-entity')

    def __init__(self, name: str, encoding_type: str, batch_size: int = 8):
        self.name = name
        assert encoding_type in SimilarityModel.ENCODING_TYPES, \
            'Model output representation must be "abstract", "sentence" or "sentence-entity"'
        self.encoding_type = encoding_type
        self.batch_size = batch_size
        self.cache = None

    @abstractmethod
    def encode(self, batch_papers: List[Dict]):
        raise NotImplementedError()

    @abstractmethod
    def get_similarity(self, x: Union[Tensor, np.ndarray], y: Union[Tensor, np.ndarray]):
        raise NotImplementedError()

    def set_encodings_cache(self, cache_filename: str):
        self.cache = EncodingsCache(cache_filename)

    def cache_encodings(self, batch_pids: List[str], batch_papers: List[dict]):
        assert self.cache is not None, "Cannot cache encodings, cache is not set"
        encodings = self.encode(batch_papers)
        for pid, enc in zip(batch_pids, encodings):
            self.cache.create_dataset(name=pid, data=enc)
        return encodings

    def get_encoding(self, pids: List[str], dataset) -> Dict:
        uncached = [p for p in pids if p not in self.cache] if self.cache is not None else list(pids)
        out = {p: torch.from_numpy(np.array(self.cache.get(p))) for p in set(pids).difference(uncached)}
        for batch_pids, batch_papers in batchify({p: dataset.get(p) for p in uncached}, self.batch_size):
            encs = self.cache_encodings(batch_pids, batch_papers) if self.cache is not None \
                else self.encode(batch_papers)
            out.update({p: encs[i] for i, p in enumerate(batch_pids)})
        return out

    def get_faceted_encoding(self, unfaceted_encoding, facet: str, input_data: Dict):
        """Keep only the sentence (and entity) rows whose predicted facet label matches (utils/models.py:127-163)."""
        if self.encoding_type == 'abstract':
            return unfaceted_encoding
        labels = ['background' if lab == 'objective_label' else lab[:-len('_label')] for lab in input_data['FACETS']]
        sent_ids = [i for i, k in enumerate(labels) if facet == k]
        if self.encoding_type == 'sentence':
            return unfaceted_encoding[sent_ids]
        ner_ids, cursor = [], len(labels)
        for i, sent_ners in enumerate(input_data['ENTITIES']):
            if i in sent_ids:
                ner_ids += list(range(cursor, cursor + len(sent_ners)))
            cursor += len(sent_ners)
        return unfaceted_encoding[sent_ids + ner_ids]

    def __del__(self):
        if getattr(self, 'cache', None) is not None:
            self.cache.close()


def _pad_tensors(encs, device, smax, D):
    """[S_j, D] tensors -> zero-padded fp32 [N, smax, D] on ``device`` with torch ops only (encodings already on the GPU)."""
    moved = [e.detach().to(device=device, dtype=torch.float32) for e in encs]
    padded = torch.nn.utils.rnn.pad_sequence(moved, batch_first=True)
    out = torch.zeros((len(encs), smax, D), dtype=torch.float32, device=device)
    out[:, :padded.shape[1]] = padded
    return out


def pack_pool(cand_encs, device, max_sents=None):
    """list of [S_j, D] encodings (torch / numpy) -> (fp32 [N, Smax, D] zero padded on ``device``, int32 lens [N]).

    Encodings that already live on the GPU are padded there; host encodings are gathered into one (pinned) staging
    buffer by ``asp_pack_pool`` -- a few memcpy threads instead of the per-candidate Python loop of ``caching_score``
    (disent_models.py:274-290), which at 1 000 candidates costs 13-17 ms (3-4 ms here) in front of a 0.05 ms kernel."""
    n = len(cand_encs)
    lens = [int(e.shape[0]) for e in cand_encs]
    smax = max_sents or max(lens)
    D = int(cand_encs[0].shape[1])
    lens_t = torch.tensor(lens, dtype=torch.int32)
    if all(isinstance(e, Tensor) and e.is_cuda for e in cand_encs):
        return _pad_tensors(cand_encs, device, smax, D), lens_t.to(device, non_blocking=True)
    keep, ptrs = [], np.empty(n, dtype=np.uintp)  # `keep` holds the converted copies alive until the gather is done
    for j, e in enumerate(cand_encs):
        if isinstance(e, Tensor):
            t = e.detach()
            if t.device.type != 'cpu' or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device='cpu', dtype=torch.float32).contiguous()
                keep.append(t)
            ptrs[j] = t.data_ptr()
        else:
            a = e if (isinstance(e, np.ndarray) and e.dtype == np.float32 and e.flags.c_contiguous) else \
                np.ascontiguousarray(e, dtype=np.float32)
            if a is not e:
                keep.append(a)
            ptrs[j] = a.ctypes.data
        if tuple(cand_encs[j].shape) != (lens[j], D):
            raise ValueError(f"pack_pool: encoding {j} has shape {tuple(cand_encs[j].shape)}, expected ({lens[j]}, {D})")
    host = torch.empty((n, smax, D), dtype=torch.float32, pin_memory=torch.cuda.is_available())
    _abi.check(_abi.lib().asp_pack_pool(ptrs.ctypes.data, lens_t.data_ptr(), n, smax, D, host.data_ptr(), 8), "asp_pack_pool")
    del keep
    return host.to(device, non_blocking=True), lens_t.to(device, non_blocking=True)


class ResidentCorpus:
    """All encodings of a dataset packed ONCE into HBM ([N, Smax, D] fp32 + lengths), pools addressed by paper id.

    The reference keeps per-paper encodings in an h5py / joblib cache and re-packs every pool on the host
    (utils/models.py:68-124, disent_models.py:274-290); here a query's pool is an int32 index list and
    ``asp_ot_score_indexed`` walks it -- no gather, no host copy per query.  180 GB hold 5.8 M ten-sentence documents.
    """

    def __init__(self, pid2enc: Dict, device=None, max_sents=None):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.pids = list(pid2enc)
        self.row = {p: i for i, p in enumerate(self.pids)}
        self.reps, self.lens = pack_pool([pid2enc[p] for p in self.pids], self.device, max_sents=max_sents)

    def indices(self, pids):
        return torch.tensor([self.row[p] for p in pids], dtype=torch.int32).to(self.device, non_blocking=True)

    def score_pool(self, query_pid, cand_pids, model_hparams=None, diameter=None, score_aggregation='l2wasserstein'):
        """Similarities (higher == closer) of paper ``query_pid`` against ``cand_pids``: fp32 CPU tensor [len(cand_pids)].
        otAspire: -OT_eps with the schedule from ``diameter`` / ``geoml_diameter`` (default: the bounding box of the whole
        corpus, computed once); ``score_aggregation='l2max'``: tsAspire."""
        hp = dict(model_hparams or {})
        qi = self.row[query_pid]
        q = self.reps[qi:qi + 1]
        q_lens = self.lens[qi:qi + 1]
        idx = self.indices(cand_pids)
        if score_aggregation == 'l2max':
            c = self.reps.index_select(0, idx.long())
            best, _i, _ = l2max_scores(q.contiguous(), q_lens.contiguous(), c, self.lens.index_select(0, idx.long()),
                                       broadcast_query=True)
            return best.cpu()
        if diameter is None:
            diameter = hp.get('geoml_diameter')
        if diameter is None:
            if not hasattr(self, "_diameter"):
                self._diameter = bbox_diameter(self.reps, self.reps)
            diameter = self._diameter
        eps = epsilon_schedule(diameter, hp.get('geoml_blur', 0.05), hp.get('geoml_scaling', 0.9))
        res = ot_scores(q.contiguous(), q_lens.contiguous(), self.reps, self.lens, eps, temp=hp.get('sent_sm_temp', 1.0),
                        want=("dual",), q_group=max(len(cand_pids), 1), c_index=idx)
        return (-res["dual"]).cpu()


_HOST_PIPE = {}


def _host_pipeline(dev, chunk_pairs, sc, d):
    """Per-device ping-pong staging buffers + copy stream for score_pools_host (created once, reused)."""
    key = (dev.index, chunk_pairs, sc, d)
    st = _HOST_PIPE.get(key)
    if st is None:
        st = {"bufs": [torch.empty((chunk_pairs, sc, d), dtype=torch.float32, device=dev) for _ in range(2)],
              "copy": torch.cuda.Stream(device=dev),
              "ready": [torch.cuda.Event() for _ in range(2)], "free": [torch.cuda.Event() for _ in range(2)]}
        _HOST_PIPE.clear()  # keep one configuration alive
        _HOST_PIPE[key] = st
    return st


def score_pools_host(queries, q_lens, cands, cand_lens, pool_size, diameter=None, model_hparams=None,
                     score_aggregation='l2wasserstein', chunk_queries=8):
    """Several queries, each against its own pool, all held in HOST memory (the shape of an evaluation run whose
    encodings cache lives on the host, utils/models.py:112-114).

    queries [NQ,Sq,D]; cands [NQ*pool_size,Sc,D] zero padded, candidates of query i at rows i*pool_size...
    (pin both for full PCIe rate); q_lens int32 [NQ]; cand_lens int32 [NQ*pool_size].
    The pools are streamed to the GPU in chunks of ``chunk_queries`` queries on a copy stream, double buffered, so
    the H2D copy of chunk k+1 overlaps the scoring of chunk k.  Returns {'scores': float32 pinned CPU tensor
    (higher == closer: -OT_eps, or max -dist for 'l2max'), 'device_scores': the same on the GPU}; the call returns
    after the scores have landed on the host.
    """
    hp = dict(model_hparams or {})
    dev = torch.device("cuda", torch.cuda.current_device())
    NQ, NP = queries.shape[0], cands.shape[0]
    assert NP <= NQ * pool_size and cands.dim() == 3 and queries.dim() == 3
    main = torch.cuda.current_stream(dev)
    q = queries.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
    ql = q_lens.to(dev, dtype=torch.int32, non_blocking=True)
    cl = cand_lens.to(dev, dtype=torch.int32, non_blocking=True)
    if score_aggregation != 'l2max':
        if diameter is None:
            diameter = hp.get('geoml_diameter')
        if diameter is None:
            raise ValueError("score_pools_host streams the pools and cannot derive geomloss' bounding-box diameter; "
                             "pass diameter= (or model_hparams['geoml_diameter'])")
        eps = epsilon_schedule(diameter, hp.get('geoml_blur', 0.05), hp.get('geoml_scaling', 0.9))
    chunk_queries = max(1, min(chunk_queries, NQ))
    chunk_pairs = chunk_queries * pool_size
    st = _host_pipeline(dev, chunk_pairs, cands.shape[1], cands.shape[2])
    sims = torch.empty(NP, dtype=torch.float32, device=dev)
    n_chunks = -(-NP // chunk_pairs)
    for k in range(n_chunks):
        lo, hi = k * chunk_pairs, min((k + 1) * chunk_pairs, NP)
        buf = st["bufs"][k % 2][:hi - lo]
        with torch.cuda.stream(st["copy"]):
            if k >= 2:
                st["copy"].wait_event(st["free"][k % 2])   # the kernel that read this buffer two chunks ago is done
            buf.copy_(cands[lo:hi], non_blocking=True)
            st["ready"][k % 2].record(st["copy"])
        main.wait_event(st["ready"][k % 2])
        q0 = lo // pool_size
        nqk = -(-(hi - lo) // pool_size)
        qk, qlk = q[q0:q0 + nqk], ql[q0:q0 + nqk]
        if score_aggregation == 'l2max':
            best = l2max_groups(qk, qlk, buf, cl[lo:hi], pool_size)
            sims[lo:hi] = best
        else:
            ot_scores(qk, qlk, buf, cl[lo:hi], eps, temp=hp.get('sent_sm_temp', 1.0), q_group=pool_size,
                      out={"dual": sims[lo:hi]})
        st["free"][k % 2].record(main)
    if score_aggregation != 'l2max':
        sims.neg_()
    host = torch.empty(sims.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(sims, non_blocking=True)
    main.synchronize()
    return {'scores': host, 'device_scores': sims}


def l2max_groups(q, q_lens, c, c_lens, pool_size):
    """tsAspire scores of several query pools: one launch per query (the l2max kernel takes one broadcast query)."""
    out = torch.empty(c.shape[0], dtype=torch.float32, device=c.device)
    for i in range(q.shape[0]):
        lo, hi = i * pool_size, min((i + 1) * pool_size, c.shape[0])
        if lo >= hi:
            break
        out[lo:hi] = l2max_scores(q[i:i + 1], q_lens[i:i + 1], c[lo:hi], c_lens[lo:hi], broadcast_query=True)[0]
    return out


def score_pool_tensors(query, cands, cand_lens, diameter=None, model_hparams=None, score_aggregation='l2wasserstein'):
    """One query against a packed pool held in HOST memory: H2D, score on the GPU, D2H of the scores.

    query [Sq,D] or [1,Sq,D]; cands [N,Sc,D] (zero padded; pin it for full PCIe rate); cand_lens int32 [N].
    Returns {'scores': float32 CPU [N] (higher == closer: -OT_eps, or max -dist for 'l2max'),
             'device_scores': the same on the GPU}.  The call returns after the scores have landed on the host.
    """
    hp = dict(model_hparams or {})
    dev = torch.device("cuda", torch.cuda.current_device())
    q = query if query.dim() == 3 else query[None]
    q = q.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
    c = cands.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
    c_lens = cand_lens.to(dev, dtype=torch.int32, non_blocking=True)
    q_lens = torch.tensor([q.shape[1]], dtype=torch.int32, device=dev)
    if score_aggregation == 'l2max':
        sims, _idx, _ = l2max_scores(q, q_lens, c, c_lens, broadcast_query=True)
    else:
        if diameter is None:
            diameter = hp.get('geoml_diameter') or bbox_diameter(q, c)
        eps = epsilon_schedule(diameter, hp.get('geoml_blur', 0.05), hp.get('geoml_scaling', 0.9))
        sims = -ot_scores(q, q_lens, c, c_lens, eps, temp=hp.get('sent_sm_temp', 1.0), broadcast_query=True)["dual"]
    host = torch.empty(sims.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(sims, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return {'scores': host, 'device_scores': sims}


class AspireModel(SimilarityModel):
    """otAspire (``aspire_compsci`` / ``aspire_biomed``); ``score_aggregation='l2max'`` gives tsAspire scoring."""

    MODEL_PATHS = {
        'compsci': 'allenai/aspire-contextualsentence-multim-compsci',
        'biomed': 'allenai/aspire-contextualsentence-multim-biomed',
    }

    def __init__(self, score_aggregation='l2wasserstein', model_hparams=None, **kwargs):
        super(AspireModel, self).__init__(**kwargs)
        from transformers import AutoTokenizer
        model_path = AspireModel.MODEL_PATHS[self.name.split('_')[-1]]
        self.model = AspireConSent(model_path)
        self.model.eval()
        self.tokenizer = AutoTokenizer.from_pretrained(model_path)
        self.score_aggregation = score_aggregation
        self.model_hparams = dict(model_hparams or {})

    def get_similarity(self, x, y):
        """-OT_eps between two [S,768] encodings, one pair per call (utils/models.py:190-197)."""
        x, y = torch.as_tensor(x), torch.as_tensor(y)
        xt = rep_len_tup(embed=x[None, :].permute(0, 2, 1), abs_lens=[len(x)])
        yt = rep_len_tup(embed=y[None, :].permute(0, 2, 1), abs_lens=[len(y)])
        if self.score_aggregation == 'l2max':
            from .distances import allpair_masked_dist_l2max
            return allpair_masked_dist_l2max(query=xt, cand=yt, return_pair_sims=True)[0].item()
        ot_dist = AllPairMaskedWasserstein(self.model_hparams).compute_distance(query=xt, cand=yt).item()
        return -ot_dist

    def encode(self, batch_papers: List[Dict]):
        # same batch as prepare_abstracts (utils/models.py:200), built by the library's host code when the tokenizer is
        # the plain BERT pipeline; the sentence spans arrive as the (start, end) table the pooling kernel consumes
        bert_batch, abs_lens, sent_token_idxs = prepare_abstracts_native(batch_abs=batch_papers,
                                                                         pt_lm_tokenizer=self.tokenizer)
        with torch.no_grad():
            _, batch_reps_sent = self.model.forward(bert_batch=bert_batch, abs_lens=abs_lens,
                                                    sent_tok_idxs=sent_token_idxs)
        return [batch_reps_sent[i, :abs_lens[i]] for i in range(len(abs_lens))]

    def score_pool(self, query_enc, cand_encs, diameter=None, return_primal=False):
        """Similarities (higher == closer) of ONE query against a pool, one launch on the current CUDA device.

        otAspire: -OT_eps (dual), or sum P*(-C) with ``return_primal`` (the caching_score quantity).  The epsilon
        schedule comes from ``diameter`` if given, else from the bounding box of this call's points
        (query + zero-padded pool), which is what geomloss sees in caching_score (disent_models.py:274-297).
        Returns a float32 CPU tensor [N].
        """
        dev = torch.device("cuda", torch.cuda.current_device())
        c, c_lens = pack_pool(cand_encs, dev)
        q = torch.as_tensor(np.asarray(query_enc) if not isinstance(query_enc, Tensor) else query_enc,
                            dtype=torch.float32).to(dev)[None].contiguous()
        q_lens = torch.tensor([q.shape[1]], dtype=torch.int32, device=dev)
        if self.score_aggregation == 'l2max':
            best, _idx, _ = l2max_scores(q, q_lens, c, c_lens, broadcast_query=True)
            return best.cpu()
        hp = self.model_hparams
        blur, scaling = hp.get('geoml_blur', 0.05), hp.get('geoml_scaling', 0.9)
        if diameter is None:
            diameter = hp.get('geoml_diameter') or bbox_diameter(q, c)
        eps = epsilon_schedule(diameter, blur, scaling)
        want = "primal" if return_primal else "dual"
        res = ot_scores(q, q_lens, c, c_lens, eps, temp=hp.get('sent_sm_temp', 1.0), want=(want,),
                        broadcast_query=True)
        return (res[want] if return_primal else -res[want]).cpu()


class AspireNER(AspireModel):
    """otAspire with input augmentation: the entities extracted from the abstract's sentences are appended to it as
    extra "sentences" (utils/models.py:211-233); everything downstream is the plain AspireModel path."""

    def encode(self, batch_papers: List[Dict]):
        assert 'ENTITIES' in batch_papers[0], 'No NER data for input. Please run NER/extract_entity.py and' \
                                             ' place result in {dataset_dir}/{dataset_name}-ner.jsonl'
        return super(AspireNER, self).encode(self._append_entities(batch_papers))

    @staticmethod
    def _append_entities(batch_papers):
        out = []
        for sample in batch_papers:
            entities = [e for sent_entities in sample['ENTITIES'] for e in sent_entities]
            out.append({'TITLE': sample['TITLE'], 'ABSTRACT': sample['ABSTRACT'] + entities})
        return out


class AspireContextNER(AspireModel):
    """otAspire where every entity is represented by the mean of the contextual token vectors of its span inside the
    sentence it occurs in (utils/models.py:607-734 with AspireConSenContextual :413-507).  The sentence spans and the
    entity spans go through the SAME span mean-pool kernel (K1) over one encoder forward; a paper's encoding is its
    sentence rows followed by the rows of the entities that could be located in the tokenisation."""

    def encode(self, input_data: List[Dict]):
        from .consent import span_mean_pool, spans_from_token_idxs
        bert_batch, abs_lens, sent_token_idxs, ner_token_idxs = self._preprocess_input(input_data)
        # one list of spans per paper: sentences first, then the entities that were found (contiguous token ranges)
        merged, kept = [], []
        for sents, ners in zip(sent_token_idxs, ner_token_idxs):
            found = [n for n in ners if len(n) > 0]
            merged.append(list(sents) + found)
            kept.append(len(found))
        dev = self.model._device()
        tokid_tt, seg_tt, attnmask_tt = (bert_batch[k].to(dev, non_blocking=True)
                                         for k in ('tokid_tt', 'seg_tt', 'attnmask_tt'))
        with torch.no_grad():
            hidden = self.model.encode_hidden(tokid_tt, seg_tt, attnmask_tt, seq_lens=bert_batch.get('seq_lens'))
            spans = spans_from_token_idxs(merged, max(len(m) for m in merged)).to(dev, non_blocking=True)
            _, reps = span_mean_pool(hidden, spans)
        out_dev = bert_batch['tokid_tt'].device
        return [reps[i, :abs_lens[i] + kept[i]].to(out_dev) for i in range(len(abs_lens))]

    def _preprocess_input(self, input_data):
        bert_batch, abs_lens, sent_token_idxs = prepare_abstracts(batch_abs=input_data, pt_lm_tokenizer=self.tokenizer)
        return bert_batch, abs_lens, sent_token_idxs, self._get_ner_token_idxs(input_data, sent_token_idxs)

    def _get_ner_token_idxs(self, input_data, sent_token_idxs):
        """Token positions of every entity: the entity's word pieces are searched in its sentence's word pieces; an
        entity that is not found, or that falls (partly) behind the 500-piece truncation, gets [] (:661-682).

        The reference tokenises every sentence and every entity with its own ``tokenizer.tokenize`` call; here all of
        them go through ONE call of the library's word-piece front end when the tokenizer is the plain BERT pipeline
        (ids instead of token strings: the vocabulary is a bijection, and unknown words are the same [UNK] either way)."""
        from .consent import native_wordpiece
        wp = native_wordpiece(self.tokenizer)
        pieces_of = None
        if wp is not None:
            texts = []
            for sample, sample_sent_idxs in zip(input_data, sent_token_idxs):
                for ners, sentence, _ in zip(sample['ENTITIES'], sample['ABSTRACT'], sample_sent_idxs):
                    texts.append(sentence)
                    texts.extend(ners)
            ids, offs = wp.encode(texts)
            ids, offs, cursor = ids.tolist(), offs.tolist(), [0]

            def pieces_of(_text):  # texts are consumed in the order they were collected
                k = cursor[0]
                cursor[0] += 1
                return ids[offs[k]:offs[k + 1]]
        else:
            pieces_of = self.tokenizer.tokenize
        all_idxs = []
        for sample, sample_sent_idxs in zip(input_data, sent_token_idxs):
            sample_idxs = []
            for ners, sentence, token_idxs in zip(sample['ENTITIES'], sample['ABSTRACT'], sample_sent_idxs):
                tokens = pieces_of(sentence)
                for ner in ners:
                    rng = self.find_sublist_range(tokens, pieces_of(ner))
                    if rng and rng[-1] < len(token_idxs):
                        sample_idxs.append([token_idxs[k] for k in rng])
                    else:
                        sample_idxs.append([])
            all_idxs.append(sample_idxs)
        return all_idxs

    @staticmethod
    def find_sublist_range(suplist: List, sublist: List):
        """Positions of the first occurrence of ``sublist`` inside ``suplist`` (None when absent; :684-698)."""
        n, m = len(suplist), len(sublist)
        for i in range(n):
            if i + m <= n and suplist[i:i + m] == sublist:
                return list(range(i, i + m))
        return None

    def get_faceted_encoding(self, unfaceted_encoding, facet: str, input_data: Dict):
        """Entities without an encoding (see _get_ner_token_idxs) are dropped from the entity lists before the usual
        facet filter, so row indices line up with what ``encode`` produced (:709-734)."""
        _, _, _, ner_token_idxs = self._preprocess_input([input_data])
        is_valid = [len(x) > 0 for x in ner_token_idxs[0]]
        # NOTE: as in the reference (:721-729) the cursor into is_valid advances only past VALID entities, so the first
        # entity without an encoding also removes every entity after it; kept for result parity.
        cursor, filtered = 0, []
        for sent_ners in input_data['ENTITIES']:
            keep = []
            for entity in sent_ners:
                if is_valid[cursor]:
                    keep.append(entity)
                    cursor += 1
            filtered.append(keep)
        return super(AspireContextNER, self).get_faceted_encoding(unfaceted_encoding, facet,
                                                                  {**input_data, 'ENTITIES': filtered})


def get_model(model_name, trained_model_path=None) -> SimilarityModel:
    """Factory (utils/models.py:738-768) for the model names on the fine-grained scoring path."""
    if model_name in {'aspire_compsci', 'aspire_biomed'}:
        return AspireModel(name=model_name, encoding_type='sentence')
    if model_name in {'tsaspire_compsci', 'tsaspire_biomed'}:
        return AspireModel(name=model_name, encoding_type='sentence', score_aggregation='l2max')
    if model_name in {'aspire_ner_compsci', 'aspire_ner_biomed'}:
        return AspireNER(name=model_name, encoding_type='sentence-entity')
    if model_name in {'aspire_context_ner_compsci', 'aspire_context_ner_biomed'}:
        return AspireContextNER(name=model_name, encoding_type='sentence-entity')
    raise NotImplementedError(f"No Implementation for model {model_name}")


def _mix_scores(scores, query_encode_ret_dict, cand_encode_ret_dicts, sent_loss_prop, abs_loss_prop, dev):
    """disent_models.py:298-307: sent_loss_prop * sentence score (+ abs_loss_prop * -||cls_q - cls_c||) -> np [B]."""
    if sent_loss_prop == 1.0 and not abs_loss_prop > 0.0:
        return scores.cpu().numpy()
    B = scores.shape[0]
    if abs_loss_prop > 0.0:
        qc = torch.as_tensor(np.asarray(query_encode_ret_dict['doc_cls_reps']), dtype=torch.float32).to(dev).view(1, -1).contiguous()
        cc = torch.as_tensor(np.stack([np.asarray(d['doc_cls_reps']) for d in cand_encode_ret_dicts]),
                             dtype=torch.float32).to(dev).contiguous()
    else:  # only the scaling: the CLS term is multiplied by exactly 0 (any finite vectors do)
        qc = torch.zeros((1, 4), dtype=torch.float32, device=dev)
        cc = torch.zeros((B, 4), dtype=torch.float32, device=dev)
    scores = scores.contiguous()
    _abi.check(_abi.lib().asp_mix_cls_scores(_abi.ptr(scores), _abi.ptr(qc), max(B, 1), _abi.ptr(cc), B, int(cc.shape[1]),
                                             float(sent_loss_prop), float(max(abs_loss_prop, 0.0)), _abi.stream_of(dev)),
               "asp_mix_cls_scores")
    return scores.cpu().numpy()


def caching_score(query_encode_ret_dict, cand_encode_ret_dicts, score_agg_type='l2wasserstein',
                  model_hparams=None, sent_loss_prop=None, abs_loss_prop=None):
    """Batched 1 x B scorer with the calling convention of WordSentAlignBiEnc.caching_score
    (src/learning/facetid_models/disent_models.py:256-342).

    :param query_encode_ret_dict: {'sent_reps': np [Sq,D], 'doc_cls_reps': np [D]}
    :param cand_encode_ret_dicts: list of such dicts
    :param score_agg_type: 'l2wasserstein' | 'l2max' | 'l2lse' (scored as l2max, :294-295) | 'l2top2' | 'l2attention'
    :param sent_loss_prop, abs_loss_prop: the mixing weights of :298-307 -- batch_scores = sent_loss_prop * sentence
        score + abs_loss_prop * (-||cls_q - cls_c||_2) when abs_loss_prop > 0.  Default: ``model_hparams``'
        'sent_loss_prop' (or 'sentsup_loss_prop' if larger, :300-303) and 'abs_loss_prop', else 1 and 0 (the values
        WordSentAlignBiEnc hard-codes, :253-254).
    :return: {'batch_scores': np [B], 'pair_scores': list of un-padded per-pair outputs}
        l2wasserstein -> pair_scores[i] = [alpha[:ql], beta[:cl], -C[:ql,:cl], plan[:ql,:cl], plan*-C[:ql,:cl]]
        l2max, l2top2 -> pair_scores[i] = -dist[:ql,:cl]
        l2attention   -> pair_scores[i] = [-dist[:ql,:cl], softmax[:ql,:cl], softmax*-dist[:ql,:cl]]
    """
    hp = dict(model_hparams or {})
    if sent_loss_prop is None:
        sent_loss_prop = max(float(hp.get('sent_loss_prop', 1.0)), float(hp['sentsup_loss_prop'])) \
            if 'sentsup_loss_prop' in hp else float(hp.get('sent_loss_prop', 1.0))
    if abs_loss_prop is None:
        abs_loss_prop = float(hp.get('abs_loss_prop', 0.0))
    dev = torch.device("cuda", torch.cuda.current_device())
    c, c_lens = pack_pool([d['sent_reps'] for d in cand_encode_ret_dicts], dev)
    q = torch.as_tensor(np.asarray(query_encode_ret_dict['sent_reps']), dtype=torch.float32).to(dev)[None].contiguous()
    qn = q.shape[1]
    q_lens = torch.tensor([qn], dtype=torch.int32, device=dev)
    lens = c_lens.cpu().tolist()

    def mixed(t):
        return _mix_scores(t, query_encode_ret_dict, cand_encode_ret_dicts, sent_loss_prop, abs_loss_prop, dev)
    if score_agg_type in ('l2max', 'l2lse'):
        best, _idx, sims = l2max_scores(q, q_lens, c, c_lens, broadcast_query=True, want_pair_sims=True)
        sims = sims.cpu().numpy()
        return {'batch_scores': mixed(best), 'pair_scores': [sims[i, :qn, :n] for i, n in enumerate(lens)]}
    if score_agg_type in ('l2top2', 'l2attention'):
        from .distances import pair_heads
        B = c.shape[0]
        qrep, qlrep = q.expand(B, -1, -1).contiguous(), q_lens.expand(B).contiguous()
        if score_agg_type == 'l2top2':
            if qn * c.shape[1] < 2:
                raise RuntimeError("selected index k out of range")  # torch.topk(k=2) over one entry (pair_distances.py:336)
            res = pair_heads(qrep, qlrep, c, c_lens, want=("top2",))
            neg = (-res["dist"]).cpu().numpy()
            return {'batch_scores': mixed(res["top2"]), 'pair_scores': [neg[i, :qn, :n] for i, n in enumerate(lens)]}
        res = pair_heads(qrep, qlrep, c, c_lens, temp=hp.get('cdatt_sm_temp', 1.0), want=("att", "att_probs"), raw_pads=True)
        neg, probs = (-res["dist"]).cpu().numpy(), res["att_probs"].cpu().numpy()
        return {'batch_scores': mixed(res["att"]),
                'pair_scores': [[neg[i, :qn, :n], probs[i, :qn, :n], probs[i, :qn, :n] * neg[i, :qn, :n]]
                                for i, n in enumerate(lens)]}
    if score_agg_type != 'l2wasserstein':
        raise ValueError(f'Unknown aggregation: {score_agg_type}')
    diameter = hp.get('geoml_diameter') or bbox_diameter(q, c)
    eps = epsilon_schedule(diameter, hp.get('geoml_blur', 0.05), hp.get('geoml_scaling', 0.9))
    res = ot_scores(q, q_lens, c, c_lens, eps, temp=hp.get('sent_sm_temp', 1.0), broadcast_query=True,
                    want=("primal", "alpha", "beta", "neg_cost", "plan", "weighted"))
    host = {k: v.cpu().numpy() for k, v in res.items() if k != "primal"}
    pairs = [[host["alpha"][i, :qn], host["beta"][i, :n], host["neg_cost"][i, :qn, :n], host["plan"][i, :qn, :n],
              host["weighted"][i, :qn, :n]] for i, n in enumerate(lens)]
    return {'batch_scores': mixed(res["primal"]), 'pair_scores': pairs}
