"""CPU ORACLE — test infrastructure, not product code.

A PyTorch-CPU restatement (fp32 by default, fp64 on request) of the reference's SASRec / CAST graph:
``modules.py`` + ``models/sasrec.py`` + ``models/cast_1.py … cast_9.py`` + the TF-1.15 Adam update +
the rank rule of ``util.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product package
(``context-aware-sequential-recommendation_b200``) never does.

PARITY STATUS
  * model math: **parity unpinned** — the reference keeps its arithmetic inside TensorFlow 1.15.2
    (``requirements.txt:2``; not installable on Python 3.12, no network) and none of the reference's own
    tests (``test.py``) touch model outputs, so there are no golden tensors to pin this restatement to.
    It follows the reference source line by line (citations on every function) and TF-1.15's documented
    op semantics.
  * data side (sampler stream, negatives, time bins, eval candidates): **pinned** against the reference's own
    ``sampler.py`` / ``util.py`` executed in the build container; see ``tests/golden/make_golden.py`` and
    ``oracle/refdata.py``.

Dropout: TF's RNG stream cannot be reproduced, so every dropout site takes an explicit keep-mask through the
``drop`` callback ``drop(site:int, x:Tensor) -> Tensor`` (site numbering below).  The CUDA path regenerates its
masks from a counter-based hash; tests hand the *same* masks to the oracle.

Parameter names are role based (shared naming convention, not shared code, with the product):
  item_emb [V,H]  pos_emb [T,H]  time_emb [max_bins+1,H]  hours_emb [25,H]  days_emb [8,H]
  <tower>.<i>.ln1.{beta,gamma}  <tower>.<i>.{q,k,v}.{w,b}  <tower>.<i>.ln2.{beta,gamma}
  <tower>.<i>.{ffn1,ffn2}.{w,b}  <tower>.lnf.{beta,gamma}  mlp.{0,1}.{w,b}
with tower in {main,time,hours,days}; all dense kernels are stored [in,out] (TF layout: y = x @ w + b).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Callable, Dict, Optional

import numpy as np
import torch

NEG_FILL = float(-2 ** 32 + 1)  # modules.py:227,239

TOWER_ID = {"main": 0, "time": 1, "hours": 2, "days": 3}
SITE_EMBED = 9000      # dropout on the embedding sum          (sasrec.py:59, cast_1.py:88)
SITE_CONCAT_A = 9001   # first dropout on a concat tensor      (cast_2.py:90, cast_4.py:116, ...)
SITE_CONCAT_B = 9002   # second dropout on the widened concat  (cast_4.py:122, cast_6.py:150)
MODELS = ["cast_1", "cast_2", "cast_3", "cast_4", "cast_5", "cast_6", "cast_7", "cast_8", "cast_9",
          "sasrec", "sasrec_static"]  # main.py:28


def block_site(tower: str, block: int, k: int) -> int:
    """k: 1 = attention probabilities (modules.py:256), 2 = FFN hidden (:303), 3 = FFN output (:309)."""
    return 1000 * TOWER_ID[tower] + 10 * block + k


def no_drop(site, x):
    return x


# ReLU sites: the FFN hidden activation of a block carries its dropout site id (block_site(.., 2)); the two dense
# layers of the CAST merge MLP are SITE_MLP0 / SITE_MLP1.  RELU_HOOK(site, pre) -> bool tensor (where the unit is
# treated as active) lets a parity test resolve the derivative AT the kink: a pre-activation within fp32 noise of 0
# (|pre| ~ 1e-6 for 50-term dot products) has an implementation-defined sign, and one flipped unit moves an
# un-normalised weight gradient by O(1e-3) relative at BASELINE batch sizes.  Default (None): plain torch.relu.
SITE_MLP0, SITE_MLP1 = 9100, 9101
RELU_HOOK = None


def _relu(site, pre):
    if RELU_HOOK is None:
        return torch.relu(pre)
    return pre * RELU_HOOK(site, pre.detach()).to(pre.dtype)


# ----------------------------------------------------------------------------------------------------------
# modules.py
# ----------------------------------------------------------------------------------------------------------
def positional_encoding(dim: int, sentence_length: int) -> np.ndarray:
    """modules.py:27-37 — literal: raw column index i, true division, flat even→sin / odd→cos, f64 → f32."""
    encoded_vec = np.array([pos / np.power(10000, 2 * i / dim)
                            for pos in range(sentence_length) for i in range(dim)])
    encoded_vec[::2] = np.sin(encoded_vec[::2])
    encoded_vec[1::2] = np.cos(encoded_vec[1::2])
    return encoded_vec.reshape([sentence_length, dim]).astype(np.float32)


def normalize(x, beta, gamma, epsilon=1e-8):
    """modules.py:53-80 — biased variance, (x-mean)/(var+eps)**.5, gamma*.. + beta."""
    mean = x.mean(-1, keepdim=True)
    variance = ((x - mean) ** 2).mean(-1, keepdim=True)
    normalized = (x - mean) / ((variance + epsilon) ** 0.5)
    return gamma * normalized + beta


def embedding(ids, table, zero_pad=True, scale=True):
    """modules.py:83-164 — returns (outputs, lookup_table-as-used)."""
    num_units = table.shape[1]
    if zero_pad:
        table = torch.cat((torch.zeros(1, num_units, dtype=table.dtype), table[1:, :]), 0)
    out = table[ids.long()]
    if scale:
        out = out * (num_units ** 0.5)
    return out, table


def multihead_attention(queries, keys, wq, bq, wk, bk, wv, bv, num_heads, drop, site, causality=True):
    """modules.py:167-277 — returns (outputs, attention_weights[h*N,T_q,T_k] captured after dropout)."""
    Q = queries @ wq + bq                                           # :203
    K = keys @ wk + bk                                              # :204
    V = keys @ wv + bv                                              # :205
    Q_ = torch.cat(torch.chunk(Q, num_heads, dim=2), dim=0)         # :208-213 (h*N, T, C/h)
    K_ = torch.cat(torch.chunk(K, num_heads, dim=2), dim=0)
    V_ = torch.cat(torch.chunk(V, num_heads, dim=2), dim=0)
    outputs = Q_ @ K_.transpose(1, 2)                               # :216
    outputs = outputs / (K_.shape[-1] ** 0.5)                       # :219
    key_masks = torch.sign(torch.abs(keys.sum(-1))).detach()        # :222 (N, T_k)
    key_masks = key_masks.repeat(num_heads, 1)                      # :223
    key_masks = key_masks.unsqueeze(1).repeat(1, queries.shape[1], 1)  # :224
    paddings = torch.ones_like(outputs) * NEG_FILL                  # :227
    outputs = torch.where(key_masks == 0, paddings, outputs)        # :228
    if causality:                                                   # :232-241
        tril = torch.tril(torch.ones_like(outputs[0]))
        masks = tril.unsqueeze(0).repeat(outputs.shape[0], 1, 1)
        outputs = torch.where(masks == 0, paddings, outputs)
    outputs = torch.softmax(outputs, dim=-1)                        # :244
    query_masks = torch.sign(torch.abs(queries.sum(-1))).detach()   # :248 (N, T_q)
    query_masks = query_masks.repeat(num_heads, 1)
    query_masks = query_masks.unsqueeze(-1).repeat(1, 1, keys.shape[1])
    outputs = outputs * query_masks                                 # :253
    outputs = drop(site, outputs)                                   # :256
    attention_weights = outputs                                     # :259
    outputs = outputs @ V_                                          # :262
    outputs = torch.cat(torch.chunk(outputs, num_heads, dim=0), dim=2)  # :265
    outputs = outputs + queries                                     # :269
    return outputs, attention_weights


def feedforward(inputs, w1, b1, w2, b2, drop, site_hidden, site_out):
    """modules.py:280-318 — conv1d(k=1) == dense; ReLU; dropout; dense; dropout; += inputs."""
    outputs = _relu(site_hidden, inputs @ w1 + b1)                  # :298-300
    outputs = drop(site_hidden, outputs)                            # :301
    outputs = outputs @ w2 + b2                                     # :304-306
    outputs = drop(site_out, outputs)                               # :307
    outputs = outputs + inputs                                      # :311
    return outputs


def mlp(inputs, w0, b0, w1, b1):
    """modules.py:321-335 — both dense layers carry ReLU."""
    h = _relu(SITE_MLP0, inputs @ w0 + b0)
    h = _relu(SITE_MLP1, h @ w1 + b1)
    return h


# ----------------------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------------------
def _glorot(gen, fan_in, fan_out, shape):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1).mul(limit).float()


def model_layout(model: str, args, itemnum: int):
    """Which towers / tables / merge a registry name uses.  Returns a SimpleNamespace."""
    m = model.lower()
    assert m in MODELS, m
    L = args.num_blocks
    Lc = getattr(args, "num_context_blocks", 2)
    lay = SimpleNamespace(model=m, towers={"main": L}, tables=["item_emb"], mlp_in=0, learned_pos=False)
    if m in ("sasrec", "cast_9"):
        lay.learned_pos = True
    if m in ("cast_1", "cast_2", "cast_3", "cast_4", "cast_5", "cast_6"):
        lay.towers["time"] = L
        lay.tables.append("time_emb")
    if m in ("cast_3", "cast_4", "cast_5", "cast_6", "cast_7", "cast_8", "cast_9"):
        lay.tables += ["hours_emb", "days_emb"]
    if m == "cast_8":
        lay.towers["hours"] = L
        lay.towers["days"] = L
    if m == "cast_9":
        lay.towers["hours"] = Lc
        lay.towers["days"] = Lc
        lay.towers["time"] = Lc
        lay.tables.append("time_emb")
    lay.mlp_in = {"cast_2": 2, "cast_3": 3, "cast_4": 4, "cast_5": 3, "cast_6": 4, "cast_7": 3, "cast_8": 3,
                  "cast_9": 4}.get(m, 0)
    return lay


def init_params(model: str, args, itemnum: int, seed: int = 42) -> Dict[str, torch.Tensor]:
    """Fresh weights with TF-1.15's default distributions: glorot-uniform tables/kernels,
    zero biases, beta=0, gamma=1.  The draws are ours (TF's graph-seeded stream is not reproducible)."""
    gen = torch.Generator().manual_seed(seed)
    H, T = args.hidden_units, args.maxlen
    lay = model_layout(model, args, itemnum)
    p: Dict[str, torch.Tensor] = {}
    rows = {"item_emb": itemnum + 1, "time_emb": args.max_bins + 1, "hours_emb": 25, "days_emb": 8}
    for t in lay.tables:
        p[t] = _glorot(gen, rows[t], H, (rows[t], H))
    if lay.learned_pos:
        p["pos_emb"] = _glorot(gen, T, H, (T, H))
    for tower, nb in lay.towers.items():
        for i in range(nb):
            pre = f"{tower}.{i}."
            for ln in ("ln1", "ln2"):
                p[pre + ln + ".beta"] = torch.zeros(H)
                p[pre + ln + ".gamma"] = torch.ones(H)
            for d in ("q", "k", "v", "ffn1", "ffn2"):
                p[pre + d + ".w"] = _glorot(gen, H, H, (H, H))
                p[pre + d + ".b"] = torch.zeros(H)
        p[f"{tower}.lnf.beta"] = torch.zeros(H)
        p[f"{tower}.lnf.gamma"] = torch.ones(H)
    if lay.mlp_in:
        k = lay.mlp_in * H
        p["mlp.0.w"] = _glorot(gen, k, k, (k, k))
        p["mlp.0.b"] = torch.zeros(k)
        p["mlp.1.w"] = _glorot(gen, k, H, (k, H))
        p["mlp.1.b"] = torch.zeros(H)
    return p


# ----------------------------------------------------------------------------------------------------------
# models/sasrec.py, models/cast_*.py
# ----------------------------------------------------------------------------------------------------------
def _tower(x, p, tower, nblocks, num_heads, mask, drop, final_ln=True):
    """`L x (LN, MHA causal, LN, FFN, *= mask)` + final LN: sasrec.py:65-85, cast_1.py:42-60, …"""
    attn = None
    for i in range(nblocks):
        pre = f"{tower}.{i}."
        q = normalize(x, p[pre + "ln1.beta"], p[pre + "ln1.gamma"])                       # sasrec.py:69
        x, attn = multihead_attention(q, x, p[pre + "q.w"], p[pre + "q.b"], p[pre + "k.w"], p[pre + "k.b"],
                                      p[pre + "v.w"], p[pre + "v.b"], num_heads, drop,
                                      block_site(tower, i, 1), causality=True)             # :71-78
        z = normalize(x, p[pre + "ln2.beta"], p[pre + "ln2.gamma"])
        x = feedforward(z, p[pre + "ffn1.w"], p[pre + "ffn1.b"], p[pre + "ffn2.w"], p[pre + "ffn2.b"],
                        drop, block_site(tower, i, 2), block_site(tower, i, 3))            # :81
        x = x * mask                                                                       # :83
    if final_ln:
        x = normalize(x, p[f"{tower}.lnf.beta"], p[f"{tower}.lnf.gamma"])                  # :85
    return x, attn


def forward(model: str, p: Dict[str, torch.Tensor], args, batch, drop: Callable = no_drop):
    """Builds `seq_emb` exactly as the registry model does.  `batch` has int tensors input_seq [B,T] and,
    for CAST, time_seq / hours / days [B,T].  Returns (seq [B,T,H], item table as used, attention_weights)."""
    m = model.lower()
    H, T, h = args.hidden_units, args.maxlen, args.num_heads
    lay = model_layout(m, args, p["item_emb"].shape[0] - 1)
    dt = p["item_emb"].dtype
    ids = batch["input_seq"]
    mask = (ids != 0).to(dt).unsqueeze(-1)                                                 # sasrec.py:23
    sinus = torch.from_numpy(positional_encoding(H, T)).to(dt)

    ctx_attn = None
    hours = days = tseq = None
    # ---- INPUT-CONTEXT (hours / days) : cast_3.py:31-52, cast_8.py:31-95, cast_9.py:31-100
    if "hours_emb" in lay.tables:
        hours, _ = embedding(batch["hours"], p["hours_emb"])
        days, _ = embedding(batch["days"], p["days_emb"])
        if "hours" in lay.towers:
            hours, _ = _tower(hours, p, "hours", lay.towers["hours"], h, mask, drop)
            days, _ = _tower(days, p, "days", lay.towers["days"], h, mask, drop)
    # ---- CONTEXT (time tower) : cast_1.py:29-60, cast_9.py:104-129
    if "time" in lay.towers:
        tseq, _ = embedding(batch["time_seq"], p["time_emb"])
        tseq, ctx_attn = _tower(tseq, p, "time", lay.towers["time"], h, mask, drop)

    seq, table = embedding(ids, p["item_emb"])                                             # sasrec.py:27-36
    post_mlp = False
    if m == "sasrec":
        seq = seq + p["pos_emb"].unsqueeze(0)                                              # :39-56
        seq = drop(SITE_EMBED, seq)                                                        # :59
        seq = seq * mask                                                                   # :62
    elif m == "sasrec_static":
        seq = seq + sinus
        seq = drop(SITE_EMBED, seq)
        seq = seq * mask
    elif m == "cast_1":                                                                    # cast_1.py:86-91
        seq = seq + sinus
        seq = seq + tseq
        seq = drop(SITE_EMBED, seq)
        seq = seq * mask
    elif m == "cast_2":                                                                    # cast_2.py:85-95
        seq = (seq + sinus) * mask
        cat = drop(SITE_CONCAT_A, torch.cat([seq, tseq], dim=2))
        seq = mlp(cat, p["mlp.0.w"], p["mlp.0.b"], p["mlp.1.w"], p["mlp.1.b"])
    elif m == "cast_3":                                                                    # cast_3.py:112-124
        seq = seq + sinus
        seq = seq + tseq
        seq = seq * mask
        cat = drop(SITE_CONCAT_A, torch.cat([seq, hours, days], dim=2))
        seq = mlp(cat, p["mlp.0.w"], p["mlp.0.b"], p["mlp.1.w"], p["mlp.1.b"])
    elif m == "cast_4":                                                                    # cast_4.py:111-127
        seq = (seq + sinus) * mask
        cat = drop(SITE_CONCAT_A, torch.cat([seq, tseq], dim=2))
        cat = drop(SITE_CONCAT_B, torch.cat([cat, hours, days], dim=2))
        seq = mlp(cat, p["mlp.0.w"], p["mlp.0.b"], p["mlp.1.w"], p["mlp.1.b"])
    elif m == "cast_5":                                                                    # cast_5.py:113-114
        seq = seq + sinus
        seq = seq + tseq
        post_mlp = True
    elif m == "cast_6":                                                                    # cast_6.py:113
        seq = seq + sinus
        post_mlp = True
    elif m in ("cast_7", "cast_8"):                                                        # cast_7.py:77-87
        seq = (seq + sinus) * mask
        cat = drop(SITE_CONCAT_A, torch.cat([seq, hours, days], dim=2))
        seq = mlp(cat, p["mlp.0.w"], p["mlp.0.b"], p["mlp.1.w"], p["mlp.1.b"])
    elif m == "cast_9":                                                                    # cast_9.py:160-174
        seq = seq + p["pos_emb"].unsqueeze(0)
        cat = drop(SITE_CONCAT_A, torch.cat([seq, tseq, hours, days], dim=2))
        seq = mlp(cat, p["mlp.0.w"], p["mlp.0.b"], p["mlp.1.w"], p["mlp.1.b"])
        seq = seq * mask

    seq, main_attn = _tower(seq, p, "main", lay.towers["main"], h, mask, drop)             # sasrec.py:65-85

    if post_mlp and m == "cast_5":                                                         # cast_5.py:143-149
        cat = drop(SITE_CONCAT_A, torch.cat([seq, hours, days], dim=2))
        seq = mlp(cat, p["mlp.0.w"], p["mlp.0.b"], p["mlp.1.w"], p["mlp.1.b"])
    elif post_mlp and m == "cast_6":                                                       # cast_6.py:142-155
        cat = drop(SITE_CONCAT_A, torch.cat([seq, tseq], dim=2))
        cat = drop(SITE_CONCAT_B, torch.cat([cat, hours, days], dim=2))
        seq = mlp(cat, p["mlp.0.w"], p["mlp.0.b"], p["mlp.1.w"], p["mlp.1.b"])

    # which attention map the model object exposes as `attention_weights` (SURVEY.md §3d last column)
    attn = ctx_attn if m in ("cast_1", "cast_2", "cast_3", "cast_4", "cast_5", "cast_6") else main_attn
    return seq, table, attn


def loss_and_logits(seq, table, pos, neg, l2_emb=0.0, reg_tables=()):
    """sasrec.py:87-115 — literal BCE with +1e-24, istarget normaliser, AUC."""
    B, T, H = seq.shape
    pos = pos.reshape(B * T).long()
    neg = neg.reshape(B * T).long()
    pos_emb = table[pos]
    neg_emb = table[neg]
    seq_emb = seq.reshape(B * T, H)
    pos_logits = (pos_emb * seq_emb).sum(-1)
    neg_logits = (neg_emb * seq_emb).sum(-1)
    istarget = (pos != 0).to(seq.dtype)
    loss = (-torch.log(torch.sigmoid(pos_logits) + 1e-24) * istarget
            - torch.log(1 - torch.sigmoid(neg_logits) + 1e-24) * istarget).sum() / istarget.sum()
    if l2_emb:
        for t in reg_tables:  # tf.contrib.layers.l2_regularizer(scale)(w) = scale * sum(w**2)/2 (modules.py:153)
            loss = loss + l2_emb * (t ** 2).sum() / 2
    auc = (((torch.sign(pos_logits - neg_logits) + 1) / 2) * istarget).sum() / istarget.sum()
    return loss, auc, pos_logits, neg_logits


def test_logits(seq, table, test_item):
    """sasrec.py:93-97 — all T positions scored, last one sliced."""
    B, T, H = seq.shape
    seq_emb = seq.reshape(B * T, H)
    test_item_emb = table[torch.as_tensor(test_item).long()]
    logits = seq_emb @ test_item_emb.t()
    return logits.reshape(B, T, -1)[:, -1, :]


# ----------------------------------------------------------------------------------------------------------
# tf.train.AdamOptimizer(learning_rate=lr, beta2=0.98) — sasrec.py:120-121
# ----------------------------------------------------------------------------------------------------------
class TFAdam:
    """TF-1.15 Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v dense on every element; theta -= lr_t*m/(sqrt(v)+eps).
    beta powers are fp32 variables multiplied after each apply, as TF keeps them."""

    def __init__(self, params: Dict[str, torch.Tensor], lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.b1p = np.float32(beta1)
        self.b2p = np.float32(beta2)
        self.step_count = 0

    def step(self, params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor]):
        lr_t = np.float32(self.lr) * np.sqrt(np.float32(1) - self.b2p) / (np.float32(1) - self.b1p)
        with torch.no_grad():
            for k, w in params.items():
                g = grads[k]
                m, v = self.m[k], self.v[k]
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                w.sub_(float(lr_t) * m / (v.sqrt() + self.eps))
        self.b1p = np.float32(self.b1p * np.float32(self.b1))
        self.b2p = np.float32(self.b2p * np.float32(self.b2))
        self.step_count += 1


def train_step(model, p, opt: Optional[TFAdam], args, batch, drop=no_drop):
    """One `sess.run([auc, loss, train_op])` (main.py:212-219).  Returns (auc, loss, grads)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    seq, table, _ = forward(model, leaves, args, batch, drop)
    reg = [leaves[t] for t in ("item_emb", "pos_emb", "time_emb", "hours_emb", "days_emb") if t in leaves]
    loss, auc, _, _ = loss_and_logits(seq, table, batch["pos"], batch["neg"], getattr(args, "l2_emb", 0.0), reg)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    if opt is not None:
        opt.step(p, grads)
    return float(auc.detach()), float(loss.detach()), grads


# ----------------------------------------------------------------------------------------------------------
# util.py:318-327 — rank + metrics
# ----------------------------------------------------------------------------------------------------------
def rank_of_target(predictions: np.ndarray) -> int:
    """util.py:318-321 literal: (-logits).argsort().argsort()[0] with the installed numpy's default sort."""
    return int((-np.asarray(predictions)).argsort().argsort()[0])


def rank_counts(predictions: np.ndarray):
    """(count_greater, count_equal_excluding_self) for candidate 0 — SURVEY.md A-12 canonical rule."""
    pr = np.asarray(predictions)
    return int((pr[1:] > pr[0]).sum()), int((pr[1:] == pr[0]).sum())


def metrics_from_ranks(ranks):
    """util.py:323-339 — float64 running sums in user order."""
    NDCG = 0.0
    HT = 0.0
    valid_user = 0.0
    for rank in ranks:
        valid_user += 1
        if rank < 10:
            NDCG += 1 / np.log2(rank + 2)
            HT += 1
    return NDCG / valid_user, HT / valid_user


# ----------------------------------------------------------------------------------------------------------
# full-catalog scoring (sasrec.py:93-97 over the whole table + util.py:318-321 rank) — canonical fp32 order
# ----------------------------------------------------------------------------------------------------------
def canonical_logits(users: np.ndarray, table: np.ndarray) -> np.ndarray:
    """s[u,j] = sum_k users[u,k]*table[j,k], k ascending, product and sum each rounded to fp32 (no FMA) — the
    summation order the kernels' `canonical_dot` uses, so logits compare bit for bit."""
    u = np.asarray(users, np.float32)
    e = np.asarray(table, np.float32)
    acc = np.zeros((u.shape[0], e.shape[0]), np.float32)
    for k in range(u.shape[1]):
        acc = (acc + (u[:, k:k + 1] * e[None, :, k]).astype(np.float32)).astype(np.float32)
    return acc


def rank_full_counts(users, table, target, rated=None):
    """(count_greater, count_equal) of the target against every item in [1,V) that is not the target and not in
    rated[u] (iterable of ids per user); an id outside [1,V) as target scores 0 like the zero-pad row."""
    s = canonical_logits(users, table)
    U, V = s.shape
    gt = np.zeros(U, np.int64)
    eq = np.zeros(U, np.int64)
    for u in range(U):
        t = int(target[u])
        ts = s[u, t] if 0 < t < V else np.float32(0)
        ok = np.ones(V, bool)
        ok[0] = False
        if 0 <= t < V:
            ok[t] = False
        if rated is not None:
            r = np.asarray(list(rated[u]), np.int64)
            ok[r[(r >= 0) & (r < V)]] = False
        gt[u] = int((s[u, ok] > ts).sum())
        eq[u] = int((s[u, ok] == ts).sum())
    return gt, eq
