"""Atomic-file datasets with sensitive attributes -> the device-resident layouts of the hot path.

Replaces, for the fairness configs, the ingestion chain of the reference: `Dataset._load_feat` (recbole/data/dataset/
dataset.py:385-454: `pd.read_csv(engine='python')`), `_remap_ID_all` (920-974: `pd.factorize`, id 0 = '[PAD]'),
`_user_item_feat_preparation` (488-507: feature row index == id), `build` (1467-1514: RO shuffle = `torch.randperm`,
`split_by_ratio` grouped by user, 1362-1396, `_calcu_split_ids` 1339-1360) and `Sampler.get_used_ids`
(recbole/sampler/sampler.py:243-264) / `FullSortEvalDataLoader.__init__` (general_dataloader.py:173-207) -- with the same
id assignment, the same shuffle and the same per-user cut, so that the splits (and everything downstream) are IDENTICAL
to the reference's for a given seed (tests/test_atomic.py pins this on ml-100k against tensors exported from the
reference).  All of it is host-side numpy (vectorised; no per-row Python loops), done once per run.

Also covered, with the reference's semantics and order of application (`_data_filtering`, dataset.py:160-181): rows with a
missing user / item id dropped (624-642), `rm_dup_inter` (644-668), `val_interval` (803-821), `filter_inter_by_user_or_item`
(847-863, on by default), `user_inter_num_interval` / `item_inter_num_interval` k-core filtering (670-728, vectorised:
bincounts instead of Python Counters); `eval_args.order` RO | TO, `split` RS (grouped by user or not) | LS
(`valid_and_test`, `valid_only`, `test_only`; 1398-1450); `benchmark_filename` (pre-split `<name>.<part>.inter` files:
265-285, 1476-1479) and `normalize_field` / `normalize_all` (min-max, 577-618).  tests/test_atomic.py pins them on golden splits exported from
the reference (tests/golden/ingest_*.npz) and, in the build container, against the live reference.

File format (RecBole atomic files): TSV with a `name:type` header, types token / float / token_seq / float_seq
(sequences are not used by the fairness configs and are skipped)."""
import os

import numpy as np
import torch

from .interaction import Interaction


# what pandas.read_csv treats as missing by default -- the reference reads with those defaults (dataset.py:433-435)
NA_STRINGS = ["", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN", "<NA>", "N/A", "NA",
              "NULL", "NaN", "None", "n/a", "nan", "null"]


SEQLEN_KEY = "__seqlen__"      # read_atomic(with_seq=True): {float_seq field: true sequence lengths} next to the columns


def minmax(values, valid=None):
    """`Dataset._normalize` (dataset.py:604-618) on float64 data: (x - min) / (max - min) over the `valid` entries, 1.0
    everywhere when all of them are equal"""
    x = np.asarray(values, np.float64)
    sel = x if valid is None else x[valid]
    if sel.size == 0:
        return x
    mx, mn = sel.max(), sel.min()
    out = np.ones_like(x) if mx == mn else (x - mn) / (mx - mn)
    if valid is not None:
        out = np.where(valid, out, 0.0)
    return out


class TokenColumn:
    """A token column as int64 codes into `vocab` (object array of str, in order of first appearance in the column);
    code -1 = missing.  Tokens are identified by their TEXT ('01' != '1'); keeping codes instead of a Python string per
    row is what lets the ingestion scale with the number of rows."""
    __slots__ = ("codes", "vocab")

    def __init__(self, codes, vocab):
        self.codes, self.vocab = np.asarray(codes, np.int64), vocab

    def __len__(self):
        return len(self.codes)

    def __getitem__(self, idx):
        return TokenColumn(self.codes[idx], self.vocab)

    def isna(self):
        return self.codes < 0

    def tokens(self):
        """object array of str (NaN where missing)"""
        out = np.empty(len(self.codes), object)
        ok = self.codes >= 0
        out[ok] = self.vocab[self.codes[ok]]
        out[~ok] = np.nan
        return out

    @staticmethod
    def of(values):
        """from a TokenColumn (returned as is) or any array of str / NaN"""
        if isinstance(values, TokenColumn):
            return values
        import pandas as pd
        codes, vocab = pd.factorize(np.asarray(values, object))
        return TokenColumn(codes, np.asarray(vocab, object))


def _use_arrow():
    if os.environ.get("FAIRREC_NO_PYARROW"):
        return False
    try:
        import pyarrow.csv  # noqa: F401
        return True
    except ImportError:
        return False


def read_atomic(path, usecols=None, sep="\t", seq_sep=" ", with_seq=False):
    """-> (columns: {field: TokenColumn | np.ndarray}, types: {field: 'token' | 'float' | 'float_seq'}).  Token columns come
    back as TokenColumn, float columns as float64 (NaN where missing), float_seq columns (only with `with_seq`: the
    pretrained-embedding files of `additional_feat_suffix`) as a zero-padded float64 matrix [rows, longest sequence]
    (dataset.py:385-454).  Parsed by pyarrow's multi-threaded CSV reader when it is importable (FAIRREC_NO_PYARROW=1
    forces the pandas parser; both give the same columns, tests/test_atomic.py)."""
    with open(path, "r", encoding="utf-8") as f:
        header = f.readline().rstrip("\n").split(sep)
    names, types = zip(*[h.split(":") for h in header])
    ok = ("token", "float", "float_seq") if with_seq else ("token", "float")
    keep = [n for n, t in zip(names, types) if t in ok and (usecols is None or n in usecols)]
    kept = {n: t for n, t in zip(names, types) if n in keep}
    raw = {}
    if _use_arrow():
        import pyarrow as pa
        import pyarrow.csv as pc
        tab = pc.read_csv(path, parse_options=pc.ParseOptions(delimiter=sep),
                          read_options=pc.ReadOptions(skip_rows=1, column_names=list(names)),
                          convert_options=pc.ConvertOptions(
                              column_types={n: (pa.float64() if kept[n] == "float" else pa.string()) for n in keep},
                              include_columns=keep, strings_can_be_null=True, null_values=NA_STRINGS))
        def convert(n):
            col = tab[n].combine_chunks()
            if kept[n] == "float":
                return col.to_numpy(zero_copy_only=False).astype(np.float64, copy=False)
            if kept[n] == "token":
                d = col.dictionary_encode()                   # dictionary in order of first appearance
                return TokenColumn(d.indices.fill_null(-1).to_numpy(zero_copy_only=False),
                                   np.array(d.dictionary.to_pylist(), dtype=object))
            return col.to_pylist()

        from concurrent.futures import ThreadPoolExecutor      # arrow's kernels release the GIL: columns in parallel
        with ThreadPoolExecutor(max_workers=max(1, min(len(keep), os.cpu_count() or 1))) as pool:
            raw = dict(zip(keep, pool.map(convert, keep)))
    else:
        import pandas as pd
        df = pd.read_csv(path, sep=sep, header=0, names=list(names), usecols=keep, engine="c",
                         dtype={n: (np.float64 if kept[n] == "float" else str) for n in keep})
        for n in keep:
            v = df[n].to_numpy()
            raw[n] = v if kept[n] == "float" else (TokenColumn.of(v) if kept[n] == "token" else list(v))
    for n in keep:
        if kept[n] != "float_seq":
            continue
        rows = [np.array([float(x) for x in (v.split(seq_sep) if isinstance(v, str) else []) if x], np.float64)
                for v in raw[n]]
        mat = np.zeros((len(rows), max((len(r) for r in rows), default=0)), np.float64)
        for k, r in enumerate(rows):
            mat[k, :len(r)] = r
        raw[n] = mat
        raw.setdefault(SEQLEN_KEY, {})[n] = np.array([len(r) for r in rows], np.int64)     # the padding is not data
    return raw, kept


def unify(cols):
    """codes of several token columns in ONE code space (string hashing only over the vocabularies, not the rows)
    -> ([int64 codes per column, -1 = missing], joint vocabulary)"""
    import pandas as pd
    cols = [TokenColumn.of(c) for c in cols]
    joint, vocab = pd.factorize(np.concatenate([c.vocab for c in cols]).astype(object)) if cols else (np.zeros(0, np.int64), [])
    out, off = [], 0
    for c in cols:
        m = joint[off:off + len(c.vocab)]
        off += len(c.vocab)
        out.append(np.where(c.codes >= 0, m[np.maximum(c.codes, 0)], -1) if len(m) else np.full(len(c), -1, np.int64))
    return out, np.asarray(vocab, object)


def factorize(chunks):
    """ids by first appearance over the concatenation of `chunks` (what pd.factorize gives the reference, dataset.py:952-974),
    + 1 so that 0 = '[PAD]' (also the id of a missing token), split back -> ([ids per chunk], id -> token array)"""
    import pandas as pd
    codes, vocab = unify(chunks)
    flat = np.concatenate(codes) if codes else np.zeros(0, np.int64)
    ids = np.zeros(len(flat), np.int64)
    ok = flat >= 0
    f, uniq = pd.factorize(flat[ok])
    ids[ok] = f + 1
    out = np.split(ids, np.cumsum([len(c) for c in codes])[:-1])
    return out, np.array(["[PAD]"] + list(vocab[uniq]), dtype=object)


def calcu_split_counts(tot, ratios):
    """dataset.py:1339-1360 _calcu_split_ids, vectorised over groups: tot [n_groups] -> counts [n_groups, len(ratios)]"""
    tot = np.asarray(tot, np.int64)
    r = np.asarray(ratios, np.float64) / np.sum(ratios)
    cnt = np.stack([(r[i] * tot).astype(np.int64) for i in range(len(r))], axis=1)
    cnt[:, 0] = tot - cnt[:, 1:].sum(axis=1)
    for i in range(1, len(r)):
        frac = r[-i] * tot
        move = (cnt[:, 0] > 1) & (frac > 0) & (frac < 1)
        cnt[move, -i] += 1
        cnt[move, 0] -= 1
    return cnt


def parse_intervals(text):
    """dataset.py:748-774 "(0,1];[3,inf)" -> [(left bracket, left, right, right bracket)] (None stays None)"""
    if text is None:
        return None
    out = []
    for part in str(text).split(";"):
        part = part.strip()
        lb, rb = part[0], part[-1]
        ends = part[1:-1].split(",")
        if len(ends) != 2 or lb not in "([" or rb not in ")]":
            continue                                             # the reference warns and skips
        out.append((lb, float(ends[0]), float(ends[1]), rb))
    return out


def within_intervals(x, intervals):
    """dataset.py:776-786, elementwise over an array"""
    x = np.asarray(x, np.float64)
    res = np.ones(x.shape, bool)
    for k, (lb, lo, hi, rb) in enumerate(intervals):
        t = (x >= lo if lb == "[" else x > lo) & (x <= hi if rb == "]" else x < hi)
        res = t if k == 0 else res | t
    return res


def _take(cols, keep):
    return {k: v[keep] for k, v in cols.items()}


def _rows(tab):
    return len(next(iter(tab.values())))


def data_filtering(config, inter, user, item, types, uid_field, iid_field):
    """dataset.py:160-181 on column dicts (user / item = None when the feature file is not loaded; token columns as
    TokenColumn or arrays of str); returns the three filtered dicts, row order preserved (after `rm_dup_inter`: the
    reference's time-sorted order).  All row-sized work runs on integer codes."""
    import pandas as pd
    def tok(tab):       # token columns given as arrays of str (direct callers) become TokenColumns
        if tab is None:
            return None
        is_token = lambda k, v: isinstance(v, TokenColumn) or types.get(k) == "token" or np.asarray(v).dtype == object
        return {k: (TokenColumn.of(v) if is_token(k, v) else np.asarray(v)) for k, v in tab.items()}

    inter, user, item = tok(inter), tok(user), tok(item)
    # 1. missing ids (624-642).  Deliberate difference: the reference drops the item-less interactions by POSITION after the
    # user-less ones were already removed (`inter_feat.index[labels]`, 638-642), i.e. once a user-less row precedes them
    # it removes their neighbours instead and keeps the item-less rows (which then map to the [PAD] item); here the rows
    # that actually miss an id are the ones dropped.
    if user is not None:
        user = _take(user, ~user[uid_field].isna())
    inter = _take(inter, ~inter[uid_field].isna())
    if item is not None:
        item = _take(item, ~item[iid_field].isna())
    inter = _take(inter, ~inter[iid_field].isna())
    # 2. duplicated (user, item) pairs (644-668): pandas does it in the reference; the same two calls here (on codes), so
    # that the order among equal timestamps (sort_values' default, unstable kind) is the reference's by construction
    keep = config["rm_dup_inter"]
    if keep is not None:
        df = pd.DataFrame({"u": inter[uid_field].codes, "i": inter[iid_field].codes, "row": np.arange(_rows(inter))})
        tf = config["TIME_FIELD"] or "timestamp"
        if tf in inter:
            df["t"] = inter[tf]
            df = df.sort_values(by=["t"], ascending=True)
        df = df.drop_duplicates(subset=["u", "i"], keep=keep)
        inter = _take(inter, df["row"].to_numpy())
    # 3. value intervals (803-821), on every table that holds the field
    for field, interval in (config["val_interval"] or {}).items():
        if field not in types:
            raise ValueError(f"Field [{field}] not defined in dataset.")
        tabs = {"inter": inter, "user": user, "item": item}
        for name, tab in tabs.items():
            if tab is None or field not in tab:
                continue
            if types[field] == "float":
                ok = within_intervals(tab[field], parse_intervals(interval))
            else:
                col = tab[field]
                allowed = np.isin(col.vocab.astype(str), [str(v) for v in interval]) if len(col.vocab) else np.zeros(0, bool)
                ok = (col.codes >= 0) & allowed[np.maximum(col.codes, 0)] if len(allowed) else np.zeros(len(col), bool)
            tabs[name] = _take(tab, ok)
        inter, user, item = tabs["inter"], tabs["user"], tabs["item"]
    # joint code spaces of the id columns of the interaction table and the feature tables
    (uc, ufc), nu = _joint(inter[uid_field], user[uid_field] if user is not None else None)
    (ic, ifc), ni = _joint(inter[iid_field], item[iid_field] if item is not None else None)
    # 4. interactions of users / items absent from a loaded feature file (847-863)
    alive = np.ones(len(uc), bool)
    if config["filter_inter_by_user_or_item"] is True:
        for code, fcode, n in ((uc, ufc, nu), (ic, ifc, ni)):
            if fcode is not None:
                present = np.zeros(n, bool)
                present[fcode] = True
                alive &= present[code]
    # 5. interaction-count intervals, iterated to the fixed point (670-746)
    u_int, i_int = parse_intervals(config["user_inter_num_interval"]), parse_intervals(config["item_inter_num_interval"])
    ualive = np.ones(len(ufc), bool) if ufc is not None else None
    ialive = np.ones(len(ifc), bool) if ifc is not None else None
    if u_int is not None or i_int is not None:
        def banned(code, fcode, falive, n, interval):
            """ids present with a count outside the interval, plus feature rows whose count is below the first
            interval's left end (_get_illegal_ids_by_inter_num; a Counter drops ids whose count reached 0)"""
            cnt = np.bincount(code[alive], minlength=n) if interval else np.zeros(n, np.int64)
            ban = (cnt > 0) & ~within_intervals(cnt, interval) if interval else np.zeros(n, bool)
            if fcode is not None:
                low = np.zeros(n, bool)
                low[fcode[falive]] = True
                ban |= low & (cnt < (interval[0][1] if interval else -1))
            return ban

        while True:
            bu, bi = banned(uc, ufc, ualive, nu, u_int), banned(ic, ifc, ialive, ni, i_int)
            if not bu.any() and not bi.any():
                break
            if ufc is not None:
                ualive &= ~bu[ufc]
            if ifc is not None:
                ialive &= ~bi[ifc]
            alive &= ~(bu[uc] | bi[ic])
    if not alive.all():
        inter = _take(inter, alive)
    if ualive is not None and not ualive.all():
        user = _take(user, ualive)
    if ialive is not None and not ialive.all():
        item = _take(item, ialive)
    for name, tab in (("inter", inter), ("user", user), ("item", item)):
        if tab is not None and _rows(tab) == 0:
            raise ValueError("Some feat is empty, please check the filtering settings.")
    return inter, user, item


def _joint(a, b):
    """((codes of a, codes of b | None), size) in one code space"""
    codes, vocab = unify([a] + ([b] if b is not None else []))
    return (codes[0], codes[1] if b is not None else None), len(vocab)


def _stable_group_order(keys, n_keys):
    """np.argsort(keys, kind='stable') for int keys in [0, n_keys) as a counting sort (scipy's COO -> CSR conversion keeps
    the input order inside a row): O(n) instead of a comparison sort"""
    import scipy.sparse as sp
    n = len(keys)
    if n == 0:
        return np.zeros(0, np.int64)
    m = sp.coo_matrix((np.ones(n, np.int8), (keys, np.arange(n))), shape=(int(n_keys) + 1, n)).tocsr()
    return m.indices.astype(np.int64)


class AtomicDataset:
    """The slice of recbole.data.dataset.Dataset the hot path reads: `num`, `inter_feat`, `get_user_feature`,
    `inter_matrix`, `field2id_token`, plus `build()` -> three column dicts (train / valid / test)."""

    def __init__(self, config):
        self.config = config
        name = config["dataset"]
        root = os.path.join(config["data_path"] or "dataset", name)
        self.uid_field, self.iid_field = config["USER_ID_FIELD"], config["ITEM_ID_FIELD"]
        self.rating_field = config["RATING_FIELD"]
        load_col = config["load_col"] or {}
        bench = config["benchmark_filename"]
        self.file_size_list = None
        if bench:
            # pre-split interaction files `<name>.<part>.inter` (dataset.py:265-285): concatenated for the id remap, handed
            # back part by part by build(); the reference applies NO data filtering to them (dataset.py:150-151)
            parts = []
            for part in bench:
                path = os.path.join(root, f"{name}.{part}.inter")
                if not os.path.isfile(path):
                    raise ValueError(f"File {path} not exist.")
                tab, itypes = read_atomic(path, load_col.get("inter"))
                parts.append(tab)
            self.file_size_list = [_rows(t) for t in parts]
            inter = {}
            for f, t in itypes.items():
                if t == "token":
                    codes, vocab = unify([tab[f] for tab in parts])
                    inter[f] = TokenColumn(np.concatenate(codes), vocab)
                else:
                    inter[f] = np.concatenate([tab[f] for tab in parts])
        else:
            inter, itypes = read_atomic(os.path.join(root, name + ".inter"), load_col.get("inter"))
        user_path, item_path = os.path.join(root, name + ".user"), os.path.join(root, name + ".item")
        user, utypes = read_atomic(user_path, load_col.get("user")) if os.path.exists(user_path) and \
            (not load_col or "user" in load_col) else (None, {})
        item, mtypes = read_atomic(item_path, load_col.get("item")) if os.path.exists(item_path) and "item" in load_col \
            else (None, {})
        self.time_field = config["TIME_FIELD"] or "timestamp"
        if not bench:
            inter, user, item = data_filtering(config, inter, user, item, {**mtypes, **utypes, **itypes}, self.uid_field,
                                               self.iid_field)
        user, item = user or {}, item or {}
        # ---- additional feature files (dataset.py:329-349), e.g. the pretrained embeddings FairGo preloads
        self.extra, self.extra_ids, xtypes = {}, {}, {}
        for suf in config["additional_feat_suffix"] or []:
            path = os.path.join(root, f"{name}.{suf}")
            if not os.path.isfile(path):
                raise ValueError(f"Additional feature file [{path}] not found.")
            self.extra[suf], xt = read_atomic(path, load_col.get(suf) if load_col else None, with_seq=True)
            xtypes.update(xt)

        def alias_chunks(key):     # fields sharing the id space, in remap order (dataset.py:456-460, 894-927)
            out = []
            for field in dict.fromkeys(config[f"alias_of_{key}"] or []):
                for suf, tab in self.extra.items():
                    if field in tab:
                        out.append((field, suf, tab[field]))
            return out

        self.field2id_token = {}
        # ---- id remap: interactions first, then the feature file, then the alias fields (dataset.py:894-927)
        ua, ia = alias_chunks("user_id"), alias_chunks("item_id")
        chunks = [inter[self.uid_field]] + ([user[self.uid_field]] if self.uid_field in user else [])
        ids, self.field2id_token[self.uid_field] = factorize(chunks + [c for _, _, c in ua])
        inter_u = ids[0]
        feat_u = ids[1] if self.uid_field in user else None
        for (field, _, _), got in zip(ua, ids[len(chunks):]):
            self.extra_ids[field], self.field2id_token[field] = got, self.field2id_token[self.uid_field]
        chunks = [inter[self.iid_field]] + ([item[self.iid_field]] if self.iid_field in item else [])
        ids, self.field2id_token[self.iid_field] = factorize(chunks + [c for _, _, c in ia])
        inter_i = ids[0]
        for (field, _, _), got in zip(ia, ids[len(chunks):]):
            self.extra_ids[field], self.field2id_token[field] = got, self.field2id_token[self.iid_field]
        self.user_num, self.item_num = len(self.field2id_token[self.uid_field]), len(self.field2id_token[self.iid_field])
        cols = {self.uid_field: inter_u.astype(np.int64), self.iid_field: inter_i.astype(np.int64)}
        for f, t in itypes.items():
            if f in (self.uid_field, self.iid_field):
                continue
            cols[f] = inter[f].astype(np.float32) if t == "float" else factorize([inter[f]])[0][0]
        # ---- label by threshold (dataset.py:865-892; the rating column is kept)
        thr = config["threshold"]
        if thr:
            (field, value), = thr.items()
            cols[config["LABEL_FIELD"]] = (cols[field] >= value).astype(np.float32)
        # ---- min-max normalisation (dataset.py:577-618; after the label, on the float64 file values, then float32)
        all_types = {**mtypes, **utypes, **itypes, **xtypes}
        if thr:
            all_types[config["LABEL_FIELD"]] = "float"
        if config["normalize_field"] is not None and config["normalize_all"] is True:
            raise ValueError("Normalize_field and normalize_all can't be set at the same time.")
        if config["normalize_field"]:
            for f in config["normalize_field"]:
                if f not in all_types:
                    raise ValueError(f"Field [{f}] does not exist.")
            norm = {f for f in config["normalize_field"] if all_types[f] in ("float", "float_seq")}
        elif config["normalize_all"]:      # `float_like_fields` lists the inter / user / item sources only (dataset.py:1012-1020:
            norm = {f for f, t in all_types.items() if t == "float" and f not in xtypes}      # additional files are not in FeatureSource)
        else:
            norm = set()
        for f in norm:
            if f in cols:
                cols[f] = minmax(inter[f] if f in inter else cols[f]).astype(np.float32)
            for tab in self.extra.values():
                if f in tab and f in tab.get(SEQLEN_KEY, {}):
                    width = tab[f].shape[1]
                    tab[f] = minmax(tab[f], np.arange(width)[None, :] < tab[SEQLEN_KEY][f][:, None])
                elif f in tab and xtypes.get(f) == "float":
                    tab[f] = minmax(tab[f])
        self.inter = cols
        # ---- user features: row index == user id, row 0 = [PAD] (dataset.py:488-507); token attributes get ids by first
        # appearance in the FILE (dataset.py:927-929), float attributes stay as they are
        self.user_feat = {self.uid_field: np.arange(self.user_num, dtype=np.int64)}
        if feat_u is not None:
            for f, t in utypes.items():
                if f == self.uid_field:
                    continue
                if t == "token":
                    vals, self.field2id_token[f] = factorize([user[f]])
                    col = np.zeros(self.user_num, np.int64)
                    col[feat_u] = vals[0]
                else:     # users without a feature row (and the [PAD] row) get the column mean (_fill_nan, dataset.py:571-572)
                    col = np.full(self.user_num, np.nanmean(user[f]) if len(user[f]) else 0.0, np.float64)
                    col[feat_u] = user[f]
                    col[np.isnan(col)] = np.nanmean(user[f]) if len(user[f]) else 0.0       # empty cells of the file, too
                    if f in norm:
                        col = minmax(col)
                    col = col.astype(np.float32)
                self.user_feat[f] = col

    # ------------------------------------------------------------------ the reference's Dataset surface
    def num(self, field):
        if field == self.uid_field:
            return self.user_num
        if field == self.iid_field:
            return self.item_num
        return len(self.field2id_token[field])

    def get_preload_weight(self, field):
        """dataset.py:505-549, 1762-1775: `preload_weight: {id field: value field}` of one additional feature file ->
        float64 [num(id field), width], row = id (row 0 and ids without a row in the file stay zero)"""
        pw = self.config["preload_weight"] or {}
        if field not in pw:
            raise ValueError(f"Field [{field}] not in preload_weight")
        value = pw[field]
        tab = next((t for t in self.extra.values() if field in t and value in t), None)
        if tab is None or field not in self.extra_ids:
            raise ValueError(f"Preload id field [{field}] / value field [{value}] must come from one additional feature "
                             f"file and the id field must be an alias of the user or item id")
        vals = tab[value]
        out = np.zeros(self.num(field)) if vals.ndim == 1 else np.zeros((self.num(field), vals.shape[1]))
        out[self.extra_ids[field]] = vals
        return out

    @property
    def inter_feat(self):
        return {k: torch.from_numpy(v) for k, v in self.inter.items()}

    def get_user_feature(self):
        return Interaction({k: torch.from_numpy(v) for k, v in self.user_feat.items()})

    def inter_matrix(self, form="coo", value_field=None):
        """dataset.py:1633-1651 (of whatever interactions this object holds: the train split after build())"""
        import scipy.sparse as sp
        src = getattr(self, "_matrix_src", self.inter)
        data = src[value_field] if value_field else np.ones(len(src[self.uid_field]), np.float32)
        m = sp.coo_matrix((data, (src[self.uid_field], src[self.iid_field])), shape=(self.user_num, self.item_num))
        return m if form == "coo" else m.tocsr()

    def __len__(self):
        return len(self.inter[self.uid_field])

    # ------------------------------------------------------------------ ordering + splitting
    def build(self):
        """dataset.py:1467-1514.  eval_args.order: RO = shuffle with torch.randperm (global torch RNG, as seeded by
        init_seed; interaction.py:293-297) | TO = stable sort by the time field.  eval_args.split: {RS: [a, b, c]} with
        group_by user (per user -- users in order of first appearance in the ordered data, rows in that order -- the
        first a/(a+b+c) go to train, etc.) or none (three contiguous ranges) | {LS: valid_and_test | valid_only |
        test_only} (leave-one-out per user).  Returns three column dicts (a part may be empty)."""
        if self.file_size_list is not None:          # dataset.py:1476-1479: the benchmark files as they are, no shuffle
            edges = np.r_[0, np.cumsum(self.file_size_list)]
            splits = [{k: v[a:b] for k, v in self.inter.items()} for a, b in zip(edges[:-1], edges[1:])]
            self._matrix_src = splits[0]
            return splits
        ea = self.config["eval_args"] or {}
        order_mode = ea.get("order") or "RO"
        split = ea.get("split") or {"RS": [8, 1, 1]}
        group_by = ea.get("group_by", "user")
        if not isinstance(split, dict) or len(split) != 1:
            raise ValueError(f"The split_args [{split}] should be a dict with one key.")
        n = len(self)
        if order_mode == "RO":
            perm = torch.randperm(n).numpy()
        elif order_mode == "TO":
            if self.time_field not in self.inter:
                raise ValueError(f"[{self.time_field}] is not exist in interaction.")
            perm = np.argsort(self.inter[self.time_field], kind="stable")       # interaction.py:333-337
        else:
            raise NotImplementedError(f"The ordering_method [{order_mode}] has not been implemented.")
        take = lambda idx: {k: v[idx] for k, v in self.inter.items()}      # one gather per column and split
        mode = next(iter(split))
        if mode == "RS" and (group_by is None or str(group_by).lower() == "none"):
            cnt = calcu_split_counts([n], split["RS"])[0]
            edges = np.r_[0, np.cumsum(cnt)]
            splits = [take(perm[a:b]) for a, b in zip(edges[:-1], edges[1:])]
            self._matrix_src = splits[0]
            return splits
        if mode not in ("RS", "LS") or (mode == "RS" and group_by != "user"):
            raise NotImplementedError(f"The splitting_method [{split}] / grouping [{group_by}] has not been implemented.")
        u = self.inter[self.uid_field][perm]
        first = np.full(self.user_num, n, np.int64)
        np.minimum.at(first, u, np.arange(n))
        order = _stable_group_order(first[u], n)              # groups by first appearance, ordered rows inside
        us = u[order]
        starts = np.flatnonzero(np.r_[True, us[1:] != us[:-1]])
        lens = np.diff(np.r_[starts, n])
        rank = np.arange(n) - np.repeat(starts, lens)
        if mode == "RS":
            n_parts = len(split["RS"])
            cnt = calcu_split_counts(lens, split["RS"])
            edges = np.cumsum(cnt, axis=1)
            part = np.zeros(n, np.int8)
            for p in range(n_parts - 1):
                part += rank >= np.repeat(edges[:, p], lens)
            slot = list(range(n_parts))
        else:
            lom = split["LS"]
            if lom not in ("valid_and_test", "valid_only", "test_only"):
                raise NotImplementedError(f"The leave_one_mode [{lom}] has not been implemented.")
            leave = 2 if lom == "valid_and_test" else 1      # dataset.py:1398-1418: the last min(leave, len - 1) rows
            legal = np.minimum(leave, lens - 1)              # of a user go one each to the LAST `legal` parts
            tot, leg = np.repeat(lens, lens), np.repeat(legal, lens)
            part = np.where(rank < tot - leg, 0, leave + 1 - leg + (rank - (tot - leg)))
            n_parts = 3
            slot = [0, 1, 2] if lom == "valid_and_test" else ([0, 1, None] if lom == "valid_only" else [0, None, 1])
        splits = [take(perm[order[part == p]] if p is not None else np.zeros(0, np.int64)) for p in slot]
        self._matrix_src = splits[0]
        return splits


def used_and_positive_lists(splits, phase, uf="user_id", itf="item_id"):
    """sampler.py:243-264 + general_dataloader.py:173-207: eval users (ascending), their positives of the phase (in split
    order) and history = the items used up to and including the phase minus the phase's positives (valid: train - valid;
    test: (train + valid) - test; the subtraction only matters when a (user, item) pair occurs more than once)."""
    ev = splits[1] if phase == "valid" else splits[2]
    used = [splits[0]] + ([splits[1]] if phase == "test" else [])
    ku = np.concatenate([s[uf] for s in used])
    ki = np.concatenate([s[itf] for s in used])
    n_items = int(max(ki.max(initial=0), ev[itf].max(initial=0))) + 1
    import pandas as pd
    dup = pd.Series(ku * n_items + ki).isin(ev[uf] * n_items + ev[itf]).to_numpy()     # hash set of the (small) eval side
    if dup.any():
        ku, ki = ku[~dup], ki[~dup]

    def group(u, i):
        o = np.argsort(u, kind="stable")
        u, i = u[o], i[o]
        s = np.flatnonzero(np.r_[True, u[1:] != u[:-1]]) if len(u) else np.zeros(0, np.int64)
        return u[s], np.split(i, s[1:])

    users, pos = group(ev[uf], ev[itf])
    hu, hl = group(ku, ki)
    hmap = dict(zip(hu.tolist(), hl))
    hist = [hmap.get(int(x), np.zeros(0, np.int64)) for x in users]
    return users, hist, pos
