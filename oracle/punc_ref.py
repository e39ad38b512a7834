"""ORACLE (test infrastructure only; only tests/, smoke() and bench.py's cpu_baseline leg may import it): CPU restatement of
the CT-Transformer punctuation path the reference runs after the acoustic model.

  * `Tokenizer`  -- funasr::CTokenizer without jieba (onnxruntime/src/tokenizer.cpp: OpenYaml :136-187, String2Ids :212-224,
                    SplitChineseString :268-284, StrSplit :294-310, Tokenize :312-365).
  * `add_punc`   -- funasr::CTTransformer::AddPunc (ct-transformer.cpp:40-157): 20-token mini-sentences, the sentence-end cache,
                    the forced period at the last comma after 200 cached tokens, the tail fix-up, en-bpe symbol mapping.
  * `infer_ids`  -- CTTransformer::Infer's argmax (ct-transformer.cpp:164-203): `Argmax(row, row + CANDIDATE_NUM - 1)`, i.e. over
                    the first FIVE of the six classes (the last punc_list entry can never be chosen by the reference).
  * `forward`    -- the network behind the session: upstream FunASR's CTTransformer (Embedding -> SANMEncoder with
                    input_layer "pe" -> Linear), restated from the published architecture.  The ONNX export is NOT in
                    /root/reference: this function is PARITY UNPINNED (like oracle/paraformer_ref.py); what the reference holds of
                    it is the call-site contract -- int32 ids [1, T] + lengths [1] in, float [1, T, 6] out.

`Tokenizer` and `add_punc` ARE pinned: the reference's own tokenizer.cpp / ct-transformer.cpp, compiled in place and run over the
stand-in onnxruntime with `forward` (or any scripted network) behind the session, produce the same strings on thousands of
random texts (tests/test_punc.py live; tests/golden/punc_golden.json elsewhere)."""
import math

import numpy as np
import torch

from . import paraformer_ref as R

TOKEN_LEN = 20                 # precomp / com-define.h
CACHE_POP_TRIGGER_LIMIT = 200
NOTPUNC, COMMA, PERIOD, QUESTION, DUN = 1, 2, 3, 4, 5
CANDIDATE_NUM = 6
UNK = "<unk>"

DEFAULT_CFG = dict(vocab=272727, d_model=256, n_heads=8, d_ff=1024, n_layers=4, kernel=11, n_punc=6, ln_eps=1e-12)
PUNC_LIST = ["<unk>", "_", "，", "。", "？", "、"]


def param_shapes(cfg):
    D, Fd, K = cfg["d_model"], cfg["d_ff"], cfg["kernel"]
    out = {"embed.weight": (cfg["vocab"], D)}
    for l in range(cfg["n_layers"]):
        p = "encoder.encoders0.0" if l == 0 else "encoder.encoders.%d" % (l - 1)
        out[p + ".norm1.weight"] = (D,)
        out[p + ".norm1.bias"] = (D,)
        out[p + ".self_attn.linear_q_k_v.weight"] = (3 * D, D)
        out[p + ".self_attn.linear_q_k_v.bias"] = (3 * D,)
        out[p + ".self_attn.fsmn_block.weight"] = (D, 1, K)
        out[p + ".self_attn.linear_out.weight"] = (D, D)
        out[p + ".self_attn.linear_out.bias"] = (D,)
        out[p + ".norm2.weight"] = (D,)
        out[p + ".norm2.bias"] = (D,)
        out[p + ".feed_forward.w_1.weight"] = (Fd, D)
        out[p + ".feed_forward.w_1.bias"] = (Fd,)
        out[p + ".feed_forward.w_2.weight"] = (D, Fd)
        out[p + ".feed_forward.w_2.bias"] = (D,)
    out["encoder.after_norm.weight"] = (D,)
    out["encoder.after_norm.bias"] = (D,)
    out["decoder.weight"] = (cfg["n_punc"], D)
    out["decoder.bias"] = (cfg["n_punc"],)
    return out


def vad_mask(T, vad_pos):
    """CTTransformerOnline::VadMask (ct-transformer-online.cpp:219-233): [T, T] of ones; with 0 < vad_pos < T rows before
    vad_pos - 1 are zero from column vad_pos on."""
    m = np.ones((T, T), np.float32)
    if 0 < vad_pos < T:
        m[:max(0, vad_pos - 1), vad_pos:] = 0.0
    return m


def _fsmn_shift(v, w, K, shift):
    left = (K - 1) // 2 + shift
    x = torch.nn.functional.pad(v.t()[None], (left, K - 1 - left))
    return torch.nn.functional.conv1d(x, w, None, groups=w.shape[0])[0].t() + v


def _mha_masked(q, k, v, H, scale, mask):
    T, D = q.shape
    dk = D // H
    qh, kh, vh = (t.reshape(T, H, dk).transpose(0, 1) for t in (q, k, v))
    s = (qh @ kh.transpose(1, 2)) * scale
    if mask is not None:
        s = s.masked_fill(torch.as_tensor(mask)[None] == 0, float("-inf"))
    return (torch.softmax(s, dim=-1) @ vh).transpose(0, 1).reshape(T, D)


def forward(ids, W, cfg, emu=False, mask=None):
    """ids int [T] -> logits float32 [T, n_punc].  emu=True rounds to bf16 where the CUDA path stores bf16.  mask [T, T]
    (realtime model): the attention mask of EVERY layer -- the reference feeds its VadMask to both the vad_mask and the
    sub_masks input of the session (ct-transformer-online.cpp:163-183); cfg["sanm_shift"] shifts the FSMN window left."""
    D, H = cfg["d_model"], cfg["n_heads"]
    shift = int(cfg.get("sanm_shift", 0))
    ids = torch.as_tensor(np.asarray(ids), dtype=torch.long)
    T = ids.shape[0]
    x = W["embed.weight"][ids] * (D ** 0.5) + R.pos_enc(T, D)
    scale = (D // H) ** -0.5
    for l in range(cfg["n_layers"]):
        p = "encoder.encoders0.0" if l == 0 else "encoder.encoders.%d" % (l - 1)
        h = R._ln(x, W, p + ".norm1", cfg["ln_eps"])
        qkv = R._lin(h, W, p + ".self_attn.linear_q_k_v", emu)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        mem = _fsmn_shift(v, W[p + ".self_attn.fsmn_block.weight"], cfg["kernel"], shift)
        att = R._rb(_mha_masked(q, k, v, H, scale, mask), emu)
        x = x + R._lin(att, W, p + ".self_attn.linear_out", emu) + mem        # in_size == size: the residual applies in layer 0 too
        h2 = R._ln(x, W, p + ".norm2", cfg["ln_eps"])
        f1 = R._rb(torch.relu(R._lin(h2, W, p + ".feed_forward.w_1", emu)), emu)
        x = x + R._lin(f1, W, p + ".feed_forward.w_2", emu)
    x = R._ln(x, W, "encoder.after_norm", cfg["ln_eps"])
    return R._rb(x, emu) @ R._rb(W["decoder.weight"], emu).t() + W["decoder.bias"]


def infer_ids(logits):
    """First maximum over classes [0, CANDIDATE_NUM - 1) per row (std::max_element; commonfunc.h:105-108)."""
    lg = np.asarray(logits, dtype=np.float32)
    out = []
    for row in lg:
        best = 0
        for j in range(1, CANDIDATE_NUM - 1):
            if row[j] > row[best]:
                best = j
        out.append(best)
    return out


class Tokenizer:
    def __init__(self, tokens, punc_list=None):
        self.id2token = list(tokens)
        self.token2id = {}
        for i, t in enumerate(tokens):
            self.token2id[t] = i           # m_token2id[element] = i: the LAST duplicate wins (tokenizer.cpp:168-172)
        self.punc = list(punc_list or PUNC_LIST)

    @staticmethod
    def split_chinese(b):
        out, i = [], 0
        while i < len(b):
            ln = 1
            for j in range(6):
                if not (b[i] & (0x80 >> j)):
                    break
                ln = j + 1
            out.append(b[i:i + ln])
            i += ln
        return out

    def tokenize(self, text):
        """-> (pieces as bytes, ids).  Works on bytes like the C++ (chars with the high bit set are 'Chinese')."""
        raw = text.encode("utf-8") if isinstance(text, str) else text
        pieces = []
        if raw != b"":
            for item in raw.split(b" "):          # StrSplit keeps empty items; they contribute nothing
                eng, chn = b"", b""
                for ch in item:
                    if not (ch & 0x80):
                        if chn:
                            pieces += self.split_chinese(chn)
                            chn = b""
                        eng += bytes([ch])
                    else:
                        if eng:
                            pieces.append(eng)
                            eng = b""
                        chn += bytes([ch])
                if chn:
                    pieces += self.split_chinese(chn)
                if eng:
                    pieces.append(eng)
        ids = []
        low = []
        for pc in pieces:
            s = bytes(c + 32 if 65 <= c <= 90 else c for c in pc)      # ::tolower on every byte, IN PLACE (String2Ids takes item by reference copy)
            low.append(s)
            key = s.decode("utf-8", "surrogateescape")
            ids.append(self.token2id.get(key, self.token2id.get(UNK, 0)))
        return pieces, ids


def add_punc(text, tok: Tokenizer, infer, language="zh-cn"):
    """infer(ids list) -> punctuation id per token (what CTTransformer::Infer returns)."""
    pieces, ids = tok.tokenize(text)
    P = [p.encode("utf-8") for p in tok.punc]
    n_total = int(math.ceil(np.float32(len(ids)) / TOKEN_LEN))
    remain_ids, remain_str = [], []
    new_punc, new_str = [], []
    sent_out, punc_out = [], []
    for i in range(0, len(ids), TOKEN_LEN):
        in_ids = remain_ids + ids[i:i + TOKEN_LEN]
        in_str = remain_str + pieces[i:i + TOKEN_LEN]
        punc = list(infer(in_ids))
        cur = i // TOKEN_LEN
        if cur < n_total - 1:
            sent_end, last_comma = -1, -1
            for k in range(len(punc) - 2, 0, -1):
                if P[punc[k]] == P[PERIOD] or P[punc[k]] == P[QUESTION]:
                    sent_end = k
                    break
                if last_comma < 0 and P[punc[k]] == P[COMMA]:
                    last_comma = k
            if sent_end < 0 and len(in_str) > CACHE_POP_TRIGGER_LIMIT and last_comma > 0:
                sent_end = last_comma
                punc[sent_end] = PERIOD
            remain_str = in_str[sent_end + 1:]
            remain_ids = in_ids[sent_end + 1:]
            in_str = in_str[:sent_end + 1]
            punc = punc[:sent_end + 1]
        new_punc += punc
        with_punc = []
        for k in range(len(in_str)):
            s = in_str[k]
            if k > 0 and not (in_str[k - 1][0] & 0x80) and not (in_str[k][0] & 0x80):
                s = b" " + s
                in_str[k] = s        # the C++ modifies InputStr[i] in place, so the next comparison sees the leading space
            with_punc.append(s)
            if punc[k] != NOTPUNC:
                with_punc.append(P[punc[k]])
        new_str += with_punc
        sent_out, punc_out = list(new_str), list(new_punc)
        if cur == n_total - 1:
            last = new_str[-1]
            if last == P[COMMA] or last == P[DUN]:
                sent_out = new_str[:-1] + [P[PERIOD]]
                punc_out = new_punc[:-1] + [PERIOD]
            elif last != P[PERIOD] and last != P[QUESTION]:
                sent_out = new_str + [P[PERIOD]]
                punc_out = new_punc + [PERIOD]
    res = b"".join(sent_out)
    if language == "en-bpe":
        for zh, en in (("，", b","), ("。", b"."), ("、", b","), ("？", b"?")):
            res = res.replace(zh.encode("utf-8"), en)
    return res.decode("utf-8", "replace")


def add_punc_online(text, cache, tok: Tokenizer, infer, language="zh-cn"):
    """funasr::CTTransformerOnline::AddPunc (ct-transformer-online.cpp:40-137).  cache: list of bytes (words kept from earlier
    calls), updated IN PLACE like arr_cache.  infer(ids, vad_pos) -> class per token; vad_pos = len(cache) at call time for every
    mini-sentence of the call."""
    P = [p.encode("utf-8") for p in tok.punc]
    raw = text.encode("utf-8") if isinstance(text, str) else text
    full = b"".join(cache)
    if len(full) > 0 and not (full[-1] & 0x80) and len(raw) > 0 and not (raw[0] & 0x80):
        full += b" "
    full += raw
    pieces, ids = tok.tokenize(full)
    n_cache = len(cache)
    n_total = int(math.ceil(np.float32(len(ids)) / TOKEN_LEN))
    remain_ids, remain_str = [], []
    punc_ids, punc_strs, words = [], [], []
    for i in range(0, len(ids), TOKEN_LEN):
        in_ids = remain_ids + ids[i:i + TOKEN_LEN]
        in_str = remain_str + pieces[i:i + TOKEN_LEN]
        punc = list(infer(in_ids, n_cache))
        if i // TOKEN_LEN < n_total - 1:
            sent_end, last_comma = -1, -1
            for k in range(len(punc) - 2, 0, -1):
                if P[punc[k]] == P[PERIOD] or P[punc[k]] == P[QUESTION]:
                    sent_end = k
                    break
                if last_comma < 0 and P[punc[k]] == P[COMMA]:
                    last_comma = k
            if sent_end < 0 and len(in_str) > CACHE_POP_TRIGGER_LIMIT and last_comma > 0:
                sent_end = last_comma
                punc[sent_end] = PERIOD
            remain_str = in_str[sent_end + 1:]
            remain_ids = in_ids[sent_end + 1:]
            in_str = in_str[:sent_end + 1]
            punc = punc[:sent_end + 1]
        punc_strs += [P[c] for c in punc]
        words += in_str
        punc_ids += punc
    out, punc_out, skip = [], [], 0
    for i in range(len(words)):
        if not (words[i][0] & 0x80) and i + 1 < len(words) and not (words[i + 1][0] & 0x80):
            words[i] = words[i] + b" "
        if skip < n_cache:
            skip += 1
        else:
            out.append(words[i])
        if skip >= n_cache:
            punc_out.append(punc_strs[i])
            if punc_strs[i] != b"_":
                out.append(punc_strs[i])
    sent_end = -1
    for i in range(len(punc_strs) - 2, 0, -1):
        if punc_ids[i] == PERIOD or punc_ids[i] == QUESTION:
            sent_end = i
            break
    cache[:] = words[sent_end + 1:]
    if len(out) > 0 and out[-1] in P:
        out = out[:-1]
    return b"".join(out).decode("utf-8", "replace")
