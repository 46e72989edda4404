"""Callers of the scoring kernels, rewritten batch-first ("next" row 1 of SURVEY 8f): the ranking loops of

  * src/evaluation/evaluate.py:35-82   ``score(model, dataset, facet, scores_filename)``
  * src/pre_process/pp_gen_nearest.py:131-204  ``CachingTrainedScoringModel.predict``
  * src/learning/facetid_models/disent_models.py:344-371  ``caching_encode``

and the file formats on either side of them: ``abstracts-{ds}.jsonl``,
``test-pid2anns-{ds}[-{facet}].json`` in, ``scores[-facet].json`` / ``{qpid: [[cpid, score], ...]}`` out -- so the
reference's ``evaluate()`` / ``ranking_eval.py`` consume the results unchanged.

The reference scores one pair per Python call (evaluate.py:72-74) or 64 candidates per call (pp_gen_nearest.py:182);
here one query's whole pool is ONE kernel launch.  geomloss derives its epsilon schedule from the bounding box of the
points of a call, so the batching decides the schedule: ``mode='reference'`` reproduces the reference's calls exactly
(one get_similarity / one 64-candidate caching_score per call), ``mode='batched'`` (default) uses one schedule per
pool -- scores move by <= 3e-5 relative, rankings agree (tests/test_evaluation_gpu.py).
"""
import codecs
import collections
import json
import logging
import os

import numpy as np
import torch

from .consent import prepare_abstracts_native
from .similarity import SimilarityModel, caching_score, pack_pool


class EvalDataset:
    """Reader of the reference's evaluation files (src/evaluation/utils/datasets.py:7-128)."""

    def __init__(self, name, root_path):
        self.name = name
        self.root_path = root_path
        self.dataset = {}
        with codecs.open(os.path.join(root_path, f'abstracts-{name}.jsonl'), 'r', 'utf-8') as fh:
            for line in fh:
                if not line.strip():
                    continue
                rec = json.loads(line)
                paper = {'TITLE': rec['title'], 'ABSTRACT': rec['abstract']}
                if 'pred_labels' in rec:
                    paper['FACETS'] = rec['pred_labels']
                self.dataset[rec['paper_id']] = paper
        ner_file = os.path.join(root_path, f'{name}-ner.jsonl')
        self.ner_data = None
        if os.path.exists(ner_file):
            with codecs.open(ner_file, 'r', 'utf-8') as fh:
                self.ner_data = json.load(fh)

    def get(self, pid):
        paper = self.dataset[pid]
        return paper if self.ner_data is None else {**paper, 'ENTITIES': self.ner_data[pid]}

    def _pool_file(self, facet):
        suffix = f'-{facet}' if facet is not None else ''
        return os.path.join(self.root_path, f'test-pid2anns-{self.name}{suffix}.json')

    def get_test_pool(self, facet=None):
        with codecs.open(self._pool_file(facet), 'r', 'utf-8') as fh:
            return json.load(fh)

    def get_gold_test_data(self, facet=None):
        return {q: dict(zip(v['cands'], v['relevance_adju'])) for q, v in self.get_test_pool(facet).items()}

    def get_test_dev_split(self):
        """{qpid: 'dev'|'test'} from ``{name}-evaluation_splits.json`` (datasets.py:106-116); None for csfcube (all test)
        and when the file is absent."""
        if self.name == 'csfcube':
            return None
        fname = os.path.join(self.root_path, f'{self.name}-evaluation_splits.json')
        if not os.path.exists(fname):
            return None
        with codecs.open(fname, 'r', 'utf-8') as fp:
            return json.load(fp)

    def get_query_metadata(self):
        """{pid: {'title': ...}} of the test-pool queries: ``{name}-queries-release.csv`` (datasets.py:96-104) when it
        exists, else the titles of the abstracts file."""
        fname = os.path.join(self.root_path, f'{self.name}-queries-release.csv')
        if os.path.exists(fname):
            import pandas as pd
            md = pd.read_csv(fname, index_col='pid')
            md.index = md.index.astype(str)
            return {pid: {'title': row['title']} for pid, row in md.iterrows()}
        return {pid: {'title': p['TITLE']} for pid, p in self.dataset.items()}

    def get_threshold_grade(self):
        return 1 if self.name in {'treccovid', 'scidcite', 'scidcocite', 'scidcoread', 'scidcoview'} else 2

    def __iter__(self):
        return iter(self.dataset.items())


def rank_candidates(candidate_pids, similarities):
    """Stable descending sort (evaluate.py:76): ties keep the pool order.  Returns [(cpid, similarity), ...]."""
    sims = np.asarray(similarities, dtype=np.float64)
    order = np.argsort(-sims, kind='stable')
    return [(candidate_pids[i], float(sims[i])) for i in order]


def score(model: SimilarityModel, dataset, facet, scores_filename, mode='batched'):
    """evaluate.py:35-82 with one launch per query pool.  Writes ``{qpid: [[cpid, -similarity], ...]}`` best first."""
    assert mode in ('batched', 'reference')
    test_pool = dataset.get_test_pool(facet=facet)
    logging.info(f"Scoring {len(test_pool)} queries in {dataset.name}" + (f', facet: {facet}' if facet is not None else ''))
    results = collections.OrderedDict()
    for query_pid, query_pool in test_pool.items():
        query_encoding = model.get_encoding(pids=[query_pid], dataset=dataset)[query_pid]
        if facet is not None:
            query_encoding = model.get_faceted_encoding(query_encoding, facet, dataset.get(query_pid))
        candidate_pids = list(dict.fromkeys(query_pool['cands']))  # a pid listed twice is ranked once (dict, evaluate.py:71-76)
        candidate_encodings = model.get_encoding(pids=candidate_pids, dataset=dataset)
        if mode == 'batched' and hasattr(model, 'score_pool'):
            sims = model.score_pool(query_encoding, [candidate_encodings[c] for c in candidate_pids]).tolist()
        else:
            sims = [model.get_similarity(query_encoding, candidate_encodings[c]) for c in candidate_pids]
        results[query_pid] = [(cpid, -1 * sim) for cpid, sim in rank_candidates(candidate_pids, sims)]
    if scores_filename:
        with codecs.open(scores_filename, 'w', 'utf-8') as fh:
            json.dump(results, fh)
        logging.info(f'Wrote: {scores_filename}')
    return results


def caching_encode(model, batch_dict):
    """disent_models.py:344-371 -- encode a prepared batch and un-pad it into per-document numpy dicts.

    ``model``: an AspireConSent; ``batch_dict``: {'bert_batch', 'abs_lens', 'senttok_idxs'}.
    Returns [{'doc_cls_reps': np [D], 'sent_reps': np [num_sents, D]}, ...].
    """
    with torch.no_grad():
        cls, reps = model.forward(bert_batch=batch_dict['bert_batch'], abs_lens=batch_dict['abs_lens'],
                                  sent_tok_idxs=batch_dict['senttok_idxs'])
    cls, reps = cls.cpu().numpy(), reps.cpu().numpy()
    return [{'doc_cls_reps': cls[i], 'sent_reps': reps[i, :n]} for i, n in enumerate(batch_dict['abs_lens'])]


class CachingScoringModel:
    """pp_gen_nearest.py:90-204 ``CachingTrainedScoringModel`` on the B200 kernels.

    ``predict(query_pid, cand_pids, pid2abstract, facet)`` encodes whatever is not cached yet (batches of 32, :141)
    and scores the pool.  ``score_batch_size=64`` reproduces the reference's 64-candidate calls (and with them its
    per-call epsilon schedules, :182-202); ``None`` scores the whole pool in one launch.
    """

    def __init__(self, model, tokenizer, score_agg_type='l2wasserstein', model_hparams=None, encode_batch_size=32,
                 score_batch_size=64):
        self.model, self.tokenizer = model, tokenizer
        self.score_agg_type = score_agg_type
        self.model_hparams = dict(model_hparams or {})
        self.encode_batch_size, self.score_batch_size = encode_batch_size, score_batch_size
        self.pid2model_reps = {}

    def save_cache(self, out_fname):
        """pp_gen_nearest.py:125-129: the whole ``pid2model_reps`` dict as one gzip-compressed joblib pickle."""
        import joblib
        joblib.dump(self.pid2model_reps, out_fname, compress=('gzip', 3))

    def load_cache(self, fname):
        """Read a ``pid2model_reps`` pickle written by ``save_cache`` here or by the reference
        ({pid: {'doc_cls_reps': np [D], 'sent_reps': np [S, D]}}, disent_models.py:365-371); documents already in the
        cache keep their entries.  Returns the number of documents read."""
        import joblib
        reps = joblib.load(fname)
        for pid, d in reps.items():
            assert 'sent_reps' in d and np.asarray(d['sent_reps']).ndim == 2, f"{fname}: entry {pid} is not a cached encoding"
            self.pid2model_reps.setdefault(pid, d)
        return len(reps)

    def _encode_missing(self, pids, pid2abstract):
        missing = [p for p in pids if p not in self.pid2model_reps]
        for s in range(0, len(missing), self.encode_batch_size):
            chunk = missing[s:s + self.encode_batch_size]
            docs = [{'TITLE': pid2abstract[p]['title'], 'ABSTRACT': pid2abstract[p]['abstract']} for p in chunk]
            bert_batch, abs_lens, senttok = prepare_abstracts_native(batch_abs=docs, pt_lm_tokenizer=self.tokenizer)
            reps = caching_encode(self.model, {'bert_batch': bert_batch, 'abs_lens': abs_lens, 'senttok_idxs': senttok})
            assert len(reps) == len(chunk)
            self.pid2model_reps.update(zip(chunk, reps))

    def predict(self, query_pid, cand_pids, pid2abstract, facet='all'):
        self._encode_missing(list(cand_pids) + [query_pid], pid2abstract)
        query_rep = self.pid2model_reps[query_pid]
        if facet != 'all':
            labs = ['background_label' if lab == 'objective_label' else lab for lab in pid2abstract[query_pid]['pred_labels']]
            keep = [i for i, lab in enumerate(labs) if lab == f'{facet}_label']
            query_rep = dict(query_rep, sent_reps=query_rep['sent_reps'][keep, :])
        cands = [self.pid2model_reps[c] for c in cand_pids]
        step = self.score_batch_size or max(len(cands), 1)
        cand_scores, pair_sm = [], []
        for s in range(0, len(cands), step):
            out = caching_score(query_rep, cands[s:s + step], score_agg_type=self.score_agg_type,
                                model_hparams=self.model_hparams)
            cand_scores.extend(out['batch_scores'].tolist())
            pair_sm.extend(out['pair_scores'])
        return {'cand_scores': cand_scores, 'pair_scores': pair_sm}

    def rank_pool(self, query_pid, cand_pids, pid2abstract, facet='all'):
        """[[cpid, score], ...] sorted by score descending, stable -- the ``test-pid2pool-*-ranked.json`` entry
        (pp_gen_nearest.py:339,360-362)."""
        scores = self.predict(query_pid, cand_pids, pid2abstract, facet)['cand_scores']
        assert len(scores) == len(cand_pids)
        return [[c, s] for c, s in rank_candidates(list(cand_pids), scores)]


def sorted_relevancies(results, dataset, facet=None):
    """{qpid: [relevance of each candidate in the model's rank order]} (utils/utils.py:70-82 ``load_score_results``).
    ``results``: what ``score`` returns / writes: {qpid: [[cpid, -similarity], ...]} best first."""
    gold = dataset.get_gold_test_data(facet)
    return {q: [gold[q][cand] for cand, _ in ranked] for q, ranked in results.items()}


FACETS = ('background', 'method', 'result')  # evaluate.py:19


def evaluate(results, dataset, facet=None, results_dir=None, pr_atks=(5, 10, 20), split_of=None):
    """Per-query and aggregated ranking metrics of a scored test pool (evaluate.py:85-157).

    ``results``: what ``score`` returns / writes for ``facet``; for ``facet='all'`` a dict {facet_i: results_i} over
    FACETS (the reference loads the three facet score files, :103-106).  Returns (per_query: list of dicts, aggregated:
    list of dicts, one per (facet, split) -- plus, for 'all', one per split over every facet's queries, :147-154).
    When ``results_dir`` is given, writes ``query-evaluations[-facet].csv`` and ``aggregated-evaluations[-facet].csv``
    there with the reference's columns (utils/utils.py:62-65); aggregated values are rounded to 4 places.
    ``split_of``: {qpid: 'dev'|'test'}; default ``dataset.get_test_dev_split()`` (``{ds}-evaluation_splits.json``, none
    for csfcube => every query is 'test').
    """
    from . import metrics as M
    by_facet = results if facet == 'all' else {('unfaceted' if facet is None else facet): results}
    if split_of is None and hasattr(dataset, 'get_test_dev_split'):
        split_of = dataset.get_test_dev_split()
    titles = dataset.get_query_metadata() if hasattr(dataset, 'get_query_metadata') else None
    thr = dataset.get_threshold_grade()
    per_query = []
    for facet_key, facet_results in by_facet.items():
        for qpid, rels in sorted_relevancies(facet_results, dataset, None if facet_key == 'unfaceted' else facet_key).items():
            m = M.compute_metrics(rels, pr_atks=list(pr_atks), threshold_grade=thr)
            m['facet'] = facet_key
            m['split'] = 'test' if split_of is None else split_of[qpid]
            m['paper_id'] = qpid
            m['title'] = titles[qpid]['title'] if titles is not None and qpid in titles else dataset.get(qpid)['TITLE']
            per_query.append(m)
    metric_cols = [k for k in per_query[0] if k not in ('facet', 'split', 'paper_id', 'title')] if per_query else []

    def first_seen(values):
        return list(dict.fromkeys(values))  # pandas .unique() order, as the reference iterates (:135-137)
    aggregated = []
    for facet_key in first_seen(m['facet'] for m in per_query):
        for split in first_seen(m['split'] for m in per_query):
            rows = [m for m in per_query if m['split'] == split and m['facet'] == facet_key]
            if not rows:
                continue
            agg = {k: round(float(np.mean([r[k] for r in rows])), 4) for k in metric_cols}
            agg['facet'] = facet_key
            agg['split'] = split
            aggregated.append(agg)
    if facet == 'all':
        for split in first_seen(m['split'] for m in per_query):
            rows = [m for m in per_query if m['split'] == split]
            agg = {k: round(float(np.mean([r[k] for r in rows])), 4) for k in metric_cols}
            agg['facet'] = facet
            agg['split'] = split
            aggregated.append(agg)
    if results_dir is not None:
        import pandas as pd
        os.makedirs(results_dir, exist_ok=True)
        suffix = '' if facet is None else f'-{facet}'
        pd.DataFrame(per_query).to_csv(os.path.join(results_dir, f'query-evaluations{suffix}.csv'), index=False)
        pd.DataFrame(aggregated).to_csv(os.path.join(results_dir, f'aggregated-evaluations{suffix}.csv'), index=False)
    return per_query, aggregated


def rank_pool_sent(root_path, reps_path, dataset, data_to_read='sent', score_type='l2max', split='', rep_type=None,
                   write=True):
    """tsAspire / l2top2 re-ranking of every test pool from pre-built sentence vectors on disk -- the numpy path
    ``rank_pool_sent`` of src/pre_process/pp_gen_nearest.py:863-985, with the ``-cdist`` + per-candidate ``max`` of
    :942-961 done by one kernel launch per query pool (``asp_l2max`` / ``asp_pair_cost`` + ``asp_pair_heads``).

    Files: ``{reps_path}/pid2idx-{dataset}-sent.json`` ({'pid-i': row}), ``{reps_path}/
    {dataset}-{data_to_read}.npy`` ([rows, D]; NaNs read as 0, :905), ``{root_path}/test-pid2anns-{dataset}{split}.json``,
    ``{root_path}/abstracts-{dataset}.jsonl``.  Returns {qpid: [(cpid, -similarity), ...]} best first (ties keep pool
    order, :962) and writes it to ``{reps_path}/test-pid2pool-{dataset}{split}-{rep_type}-ranked.json`` (:981-984).
    """
    from .distances import l2max_scores, pair_heads
    if score_type not in {'l2max', 'l2top2'}:
        raise ValueError(f'Unknown score type: {score_type}')  # cosine / dot variants are not on this path
    with codecs.open(os.path.join(root_path, f'test-pid2anns-{dataset}{split}.json'), 'r', 'utf-8') as fp:
        qpid2pool = json.load(fp)
    with codecs.open(os.path.join(reps_path, f'pid2idx-{dataset}-sent.json'), 'r', 'utf-8') as fp:
        docsent2idx = json.load(fp)
    all_reps = np.nan_to_num(np.load(os.path.join(reps_path, f'{dataset}-{data_to_read}.npy')).astype(np.float32))
    n_sents = {}
    with codecs.open(os.path.join(root_path, f'abstracts-{dataset}.jsonl'), 'r', 'utf-8') as fh:
        for line in fh:
            if line.strip():
                rec = json.loads(line)
                n_sents[rec['paper_id']] = len(rec['abstract'])
    dev = torch.device("cuda", torch.cuda.current_device())
    ranked = collections.OrderedDict()
    for qpid, pool in qpid2pool.items():
        cand_pids = list(dict.fromkeys(pool['cands']))  # cand_sims is a dict in the reference (:947-969)
        q_rows = [docsent2idx[f'{qpid}-{i}'] for i in range(n_sents[qpid])]
        q = torch.from_numpy(all_reps[q_rows])[None].to(dev).contiguous()
        q_lens = torch.tensor([len(q_rows)], dtype=torch.int32, device=dev)
        c, c_lens = pack_pool([all_reps[[docsent2idx[f'{c_}-{i}'] for i in range(n_sents[c_])]] for c_ in cand_pids], dev)
        best, _idx, _ = l2max_scores(q, q_lens, c, c_lens, broadcast_query=True)
        sims = best
        if score_type == 'l2top2':
            qrep = q.expand(c.shape[0], -1, -1).contiguous()
            top2 = pair_heads(qrep, q_lens.expand(c.shape[0]).contiguous(), c, c_lens, want=("top2",))["top2"]
            # a single sentence pair: the numpy path sums what there is (:955-957), i.e. the one similarity
            sims = torch.where((q_lens[0] * c_lens) < 2, best, top2)
        ranked[qpid] = [(cpid, -1 * s) for cpid, s in rank_candidates(cand_pids, sims.cpu().numpy())]
    if write:
        name = rep_type or os.path.basename(os.path.normpath(reps_path))
        out_fname = os.path.join(reps_path, f'test-pid2pool-{dataset}{split}-{name}-ranked.json')
        with codecs.open(out_fname, 'w', 'utf-8') as fp:
            json.dump(ranked, fp)
        logging.info(f'Wrote: {out_fname}')
    return ranked


def rank_pool_sent_treccovid(root_path, reps_path, dataset, data_to_read='sent', score_type='l2max', rep_type=None,
                             write=True, cand_chunk=16384):
    """Deep, overlapping pools (TREC-COVID: every query against most of the corpus): ``rank_pool_sent_treccovid`` of
    src/pre_process/pp_gen_nearest.py:729-860.  The reference computes ONE all-query-sentences x all-corpus-sentences
    ``-cdist`` (:788-795) and slices it per pool (:816); here every query document is scored against every document
    that occurs in any pool by the tensor-core all-pairs kernel (``asp_l2max_allpairs``, chunks of ``cand_chunk``
    documents), and a pool's scores are a gather from that [queries, documents] matrix.

    Same files, same output format and tie order as ``rank_pool_sent``; as in the reference a candidate listed twice
    in a pool is ranked once (its ``cand_sims`` is a dict, :823-848).  ``score_type``: 'l2max' / 'l2lse' (the max over
    sentence pairs; other aggregations go through ``rank_pool_sent``).
    """
    from .distances import l2max_allpairs
    if score_type not in {'l2max', 'l2lse'}:
        raise ValueError(f'Unknown score type: {score_type}')
    with codecs.open(os.path.join(root_path, f'test-pid2anns-{dataset}.json'), 'r', 'utf-8') as fp:
        qpid2pool = json.load(fp)
    with codecs.open(os.path.join(reps_path, f'pid2idx-{dataset}-sent.json'), 'r', 'utf-8') as fp:
        docsent2idx = json.load(fp)
    all_reps = np.nan_to_num(np.load(os.path.join(reps_path, f'{dataset}-{data_to_read}.npy')).astype(np.float32))
    n_sents = {}
    with codecs.open(os.path.join(root_path, f'abstracts-{dataset}.jsonl'), 'r', 'utf-8') as fh:
        for line in fh:
            if line.strip():
                rec = json.loads(line)
                n_sents[rec['paper_id']] = len(rec['abstract'])
    query_pids = list(qpid2pool)
    pools = {q: list(dict.fromkeys(qpid2pool[q]['cands'])) for q in query_pids}
    docs = list(dict.fromkeys(c for q in query_pids for c in pools[q]))  # union of the pools, first-seen order
    col_of = {c: i for i, c in enumerate(docs)}
    S = max([n_sents[p] for p in query_pids + docs] + [1])
    D = all_reps.shape[1]
    if S > 64 or D % 64 != 0:  # outside the all-pairs kernel's tiles: one launch per pool instead
        return rank_pool_sent(root_path, reps_path, dataset, data_to_read, 'l2max', '', rep_type, write)
    dev = torch.device("cuda", torch.cuda.current_device())

    def pack(pids):
        return pack_pool([all_reps[[docsent2idx[f'{p}-{i}'] for i in range(n_sents[p])]] for p in pids], dev, max_sents=S)
    q, q_lens = pack(query_pids)
    sims = torch.empty((len(query_pids), len(docs)), dtype=torch.float32, device=dev)
    for s0 in range(0, len(docs), cand_chunk):
        c, c_lens = pack(docs[s0:s0 + cand_chunk])
        sims[:, s0:s0 + c.shape[0]] = l2max_allpairs(q, q_lens, c, c_lens, want_idx=False)[0]
    sims = sims.cpu().numpy()
    ranked = collections.OrderedDict()
    for qi, qpid in enumerate(query_pids):
        cand_pids = pools[qpid]
        row = sims[qi, [col_of[c] for c in cand_pids]]
        ranked[qpid] = [(cpid, -1 * s) for cpid, s in rank_candidates(cand_pids, row)]
    if write:
        name = rep_type or os.path.basename(os.path.normpath(reps_path))
        out_fname = os.path.join(reps_path, f'test-pid2pool-{dataset}-{name}-ranked.json')
        with codecs.open(out_fname, 'w', 'utf-8') as fp:
            json.dump(ranked, fp)
        logging.info(f'Wrote: {out_fname}')
    return ranked


def encode_stream(model, tokenizer, papers, batch_size=32, workers=2):
    """Encode a list of {'TITLE', 'ABSTRACT'} papers with the host preparation of batch n+1 (tokenisation, span table:
    ``prepare_abstracts_native``, the library's multi-threaded word-piece front end, 50-100 k documents/s; the Hugging
    Face tokenizer for anything it does not cover) running on worker threads while the GPU encodes batch n -- both
    release the GIL, and at 15 k documents/s on the device the per-sentence Python protocol (0.6 k documents/s) would
    otherwise bound ``AspireModel.encode`` / ``caching_encode`` (utils/models.py:199-209, disent_models.py:344-371).

    ``model``: an AspireConSent.  Yields one fp32 CPU tensor [num_sents, 768] per paper, in input order.
    """
    from concurrent.futures import ThreadPoolExecutor
    chunks = [papers[s:s + batch_size] for s in range(0, len(papers), batch_size)]
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        pending = [pool.submit(prepare_abstracts_native, c, tokenizer) for c in chunks[:workers + 1]]
        for k in range(len(chunks)):
            bert_batch, abs_lens, spans = pending.pop(0).result()
            nxt = k + workers + 1
            if nxt < len(chunks):
                pending.append(pool.submit(prepare_abstracts_native, chunks[nxt], tokenizer))
            with torch.no_grad():
                _, reps = model.forward(bert_batch=bert_batch, abs_lens=abs_lens, sent_tok_idxs=spans)
            for i, n in enumerate(abs_lens):
                yield reps[i, :n]
