"""ORACLE (test infrastructure only) — literal Python restatement of the reference's host-side text and
timestamp post-processing on the Paraformer::Forward path.

  Vocab.vector2string_v2   Vocab::Vector2StringV2   onnxruntime/src/vocab.cpp:164-305
  Vocab.vector2string      Vocab::Vector2String     onnxruntime/src/vocab.cpp:98-104
  timestamp_onnx           TimestampOnnx            onnxruntime/src/util.cpp:838-963
  post_process             PostProcess              onnxruntime/src/util.cpp:720-836
  stitch_offline           FunOfflineInferBuffer    onnxruntime/src/funasrruntime.cpp:291-316
  fetch_dynamic            Audio::FetchDynamic      onnxruntime/src/audio.cpp:1052-1108

PARITY PINNED for Vector2StringV2 / Vector2String / TimestampOnnx / PostProcess: the reference's own vocab.cpp and util.cpp
are compiled in place into oracle/_ref/libfunasr_text_ref.so (oracle/Makefile, oracle/text_ref.py) and this file reproduces
them string for string on thousands of random inputs (tests/test_oracle_cpu.py, live when oracle/_ref is built) and on the
committed vectors generated from the compiled reference (tests/golden/text_golden.json).  stitch_offline is pinned through the
reference's own FunOfflineInferBuffer, which runs end to end over the stand-in onnxruntime (oracle/am_ref.py RefOffline,
tests/test_am_ref_cpu.py::test_reference_offline_api_end_to_end); fetch_dynamic follows audio.cpp line by line (on the reference's
CPU path it only ever forms batches of one, so its grouping cannot be observed there).  float arithmetic is done in numpy float32 where the C++ uses float, and `std::to_string(float)` is reproduced as
'%f' of the float promoted to double.
"""
import numpy as np

f32 = np.float32


def _is_chinese(ch: str) -> bool:
    """vocab.cpp:131-141 / util.cpp:708-718: a single 3-byte UTF-8 code point in [0x4E00, 0x9FFF]."""
    b = ch.encode("utf-8")
    if len(b) != 3:
        return False
    if (b[0] & 0xF0) != 0xE0 or (b[1] & 0xC0) != 0x80 or (b[2] & 0xC0) != 0x80:
        return False
    val = ((b[0] & 0x0F) << 12) | ((b[1] & 0x3F) << 6) | (b[2] & 0x3F)
    return 19968 <= val <= 40959


def _word_format(w):
    return {"i": "I", "i'm": "I'm", "i've": "I've", "i'll": "I'll"}.get(w, w)


def _blen(s):  # std::string::size() counts bytes
    return len(s.encode("utf-8"))


class Vocab:
    def __init__(self, tokens):
        self.vocab = list(tokens)
        self.last_is_complete_english = False  # vocab.h member last_is_complete_english_

    def vector2string(self, ids):
        return [self.vocab[i] for i in ids]

    def vector2string_v2(self, ids, language=""):
        words = []
        is_pre_english = False
        pre_english_len = 0
        sub_word = False
        is_combining = False
        combine = ""
        first_word_need_space = self.last_is_complete_english
        n = len(ids)
        for i in range(n):
            word = self.vocab[ids[i]]
            if word in ("<s>", "</s>", "<unk>"):
                continue
            if language == "en-bpe":
                if "▁" in word:
                    if combine != "":
                        combine = _word_format(combine)
                        if len(words) != 0:
                            combine = " " + combine
                        words.append(combine)
                    combine = word.encode("utf-8")[3:].decode("utf-8", "ignore")  # word.substr(3)
                else:
                    combine += word
                continue
            sub_word = "@@" in word
            if sub_word:
                if i < n - 1 and _is_chinese(self.vocab[ids[i + 1]]):
                    word = word[:-2] + " "
                    if is_combining:
                        combine += word
                        is_combining = False
                        word = combine
                        combine = ""
                elif i == n - 1:
                    word = word[:-2]
                    if is_combining:
                        combine += word
                        is_combining = False
                        word = combine
                        combine = ""
                    self.last_is_complete_english = False
                else:
                    combine += word[:-2]
                    is_combining = True
                    continue
            elif is_combining:
                combine += word
                is_combining = False
                word = combine
                combine = ""
            if _is_chinese(word):
                words.append(word)
                is_pre_english = False
            else:
                if (not is_pre_english) and first_word_need_space:
                    words.append(" ")
                if not is_pre_english:
                    words.append(word)
                    pre_english_len = _blen(word)
                else:
                    if pre_english_len > 1:
                        words.append(" ")
                        words.append(word)
                        pre_english_len = _blen(word)
                    else:
                        if _blen(word) > 1:
                            words.append(" ")
                        words.append(word)
                        pre_english_len = _blen(word)
                is_pre_english = True
            if i == n - 1 and (not _is_chinese(word)) and (not sub_word):
                self.last_is_complete_english = True
            else:
                self.last_is_complete_english = False
        if language == "en-bpe" and combine != "":
            combine = _word_format(combine)
            if len(words) != 0:
                combine = " " + combine
            words.append(combine)
        return "".join(words)


def to_string_f(x):
    """std::to_string(float) -> printf("%f", (double)x)."""
    return "%f" % float(f32(x))


def timestamp_onnx(us_alphas, us_cif_peak, char_list, begin_time=0.0, total_offset=-1.5):
    """util.cpp:838-963.  Returns (res_str, timestamp_vec, char_list_after, us_alphas_after); like the C++
    it mutates char_list (drops a trailing </s>) and rescales us_alphas in the num_peak mismatch branch."""
    us_alphas = [f32(a) for a in us_alphas]
    char_list = list(char_list)
    res_str = ""
    timestamp_vec = []
    if len(char_list) == 0:
        return res_str, timestamp_vec, char_list, us_alphas
    START_END_THRESHOLD = f32(5.0)
    MAX_TOKEN_DURATION = f32(30.0)
    TIME_RATE = f32(10.0 * 6 / 1000 / 3)
    cif_peak = [f32(v) for v in us_cif_peak]
    num_frames = len(cif_peak)
    if char_list[-1] == "</s>":
        char_list.pop()
    if len(char_list) == 0:
        return res_str, timestamp_vec, char_list, us_alphas
    thr = 1.0 - 1e-4  # double literal in the C++
    fire_place = [f32(i + total_offset) for i in range(num_frames) if float(cif_peak[i]) > thr]
    num_peak = len(fire_place)
    if num_peak != len(char_list) + 1:
        s = f32(0.0)
        for a in us_alphas:
            s = f32(s + a)
        scale = f32(s / f32(len(char_list) + 1))
        if scale == 0:
            return res_str, timestamp_vec, char_list, us_alphas
        cif_peak = []
        s = f32(0.0)
        for k in range(len(us_alphas)):
            us_alphas[k] = f32(us_alphas[k] / scale)
            s = f32(s + us_alphas[k])
            cif_peak.append(s)
            if float(s) >= thr:
                s = f32(float(s) - thr)
        cif_idx = len(cif_peak) - 1
        while float(s) >= thr and cif_idx >= 0:
            if float(cif_peak[cif_idx]) < thr:
                cif_peak[cif_idx] = s
                s = f32(float(s) - thr)
            cif_idx -= 1
        fire_place = [f32(i + total_offset) for i in range(num_frames) if float(cif_peak[i]) > thr]
    num_peak = len(fire_place)
    if num_peak == 0:
        return res_str, timestamp_vec, char_list, us_alphas
    timestamp_list = []
    new_char_list = []
    if fire_place[0] > START_END_THRESHOLD:
        new_char_list.append("<sil>")
        timestamp_list.append([f32(0.0), f32(fire_place[0] * TIME_RATE)])
    for i in range(num_peak - 1):
        new_char_list.append(char_list[i])
        if i == num_peak - 2 or MAX_TOKEN_DURATION < 0 or f32(fire_place[i + 1] - fire_place[i]) < MAX_TOKEN_DURATION:
            timestamp_list.append([f32(fire_place[i] * TIME_RATE), f32(fire_place[i + 1] * TIME_RATE)])
        else:
            _split = f32(fire_place[i] + MAX_TOKEN_DURATION)
            timestamp_list.append([f32(fire_place[i] * TIME_RATE), f32(_split * TIME_RATE)])
            timestamp_list.append([f32(_split * TIME_RATE), f32(fire_place[i + 1] * TIME_RATE)])
            new_char_list.append("<sil>")
    if len(timestamp_list) == 0:
        return res_str, timestamp_vec, char_list, us_alphas
    if f32(num_frames - fire_place[-1]) > START_END_THRESHOLD:
        _end = f32((float(f32(num_frames + fire_place[-1]))) / 2.0)
        timestamp_list[-1][1] = f32(_end * TIME_RATE)
        timestamp_list.append([f32(_end * TIME_RATE), f32(f32(num_frames) * TIME_RATE)])
        new_char_list.append("<sil>")
    else:
        timestamp_list[-1][1] = f32(f32(num_frames) * TIME_RATE)
    if begin_time:
        for ts in timestamp_list:
            ts[0] = f32(float(ts[0]) + float(f32(begin_time)) / 1000.0)
            ts[1] = f32(float(ts[1]) + float(f32(begin_time)) / 1000.0)
    for c, ts in zip(new_char_list, timestamp_list):
        res_str += c + " " + to_string_f(ts[0]) + " " + to_string_f(ts[1]) + ";"
    for c, ts in zip(new_char_list, timestamp_list):
        if c != "<sil>":
            timestamp_vec.append([ts[0], ts[1]])
    return res_str, timestamp_vec, char_list, us_alphas


def post_process(raw_char, timestamp_list):
    """util.cpp:720-836."""
    timestamp_merge = []
    words = []
    is_pre_english = False
    pre_english_len = 0
    is_combining = False
    combine = ""
    begin = f32(-1)
    n = len(raw_char)
    for i in range(n):
        word = raw_char[i]
        if word in ("<s>", "</s>", "<unk>"):
            continue
        sub_word = "@@" in word
        if sub_word:
            if i == n - 1 or (i < n - 1 and _is_chinese(raw_char[i + 1])):
                word = word[:-2] + " "
                if is_combining:
                    combine += word
                    is_combining = False
                    word = combine
                    combine = ""
            else:
                combine += word[:-2]
                if not is_combining:
                    begin = timestamp_list[i][0]
                is_combining = True
                continue
        elif is_combining:
            combine += word
            is_combining = False
            word = combine
            combine = ""
        if _is_chinese(word):
            words.append(word)
            timestamp_merge.append(list(timestamp_list[i]))
            is_pre_english = False
        else:
            if not is_pre_english:
                words.append(word)
            else:
                words.append(" ")
                words.append(word)
            begin = timestamp_list[i][0] if begin == -1 else begin
            timestamp_merge.append([begin, timestamp_list[i][1]])
            begin = f32(-1)
            pre_english_len = _blen(word)
            is_pre_english = True
    stamp_str = ""
    for i, ts in enumerate(timestamp_merge):
        stamp_str += to_string_f(ts[0]) + ", " + to_string_f(ts[1])
        if i != len(timestamp_merge) - 1:
            stamp_str += ","
    return "".join(words) + " | " + stamp_str


def greedy_search_text(vocab: Vocab, ids, language, us_alphas=None, us_cif_peak=None):
    """Paraformer::GreedySearch paraformer.cpp:386-408 after the argmax."""
    if us_alphas is None:
        return vocab.vector2string_v2(ids, language)
    char_list = vocab.vector2string(ids)
    raw_char = list(char_list)
    _, ts, _, _ = timestamp_onnx(us_alphas, us_cif_peak, char_list)
    return post_process(raw_char, ts)


def stitch_offline(msgs, start_times, lang):
    """funasrruntime.cpp:291-316: join per-segment messages (original time order) into text + '[[b,e],...]'."""
    text = ""
    cur_stamp = "["
    for msg, st in zip(msgs, start_times):
        if msg == "":
            continue                      # SplitStr of "" -> empty vector -> continue
        parts = msg.split(" | ")
        if lang == "en-bpe" and text != "":
            text += " "
        text += parts[0]
        if len(parts) > 1:
            stamps = [s for s in parts[1].split(",")]
            if len(stamps) > 1:
                for i in range(0, len(stamps) - 1, 2):
                    b = f32(f32(float(stamps[i])) + f32(st))
                    e = f32(f32(float(stamps[i + 1])) + f32(st))
                    cur_stamp += "[" + str(int(f32(1000) * b)) + "," + str(int(f32(1000) * e)) + "],"
    stamp = ""
    if cur_stamp != "[":
        stamp = cur_stamp[:-1] + "]"
    return text, stamp


def fetch_dynamic(lengths_sorted, batch_size, use_gpu=True, seg_sample=16):
    """Audio::FetchDynamic audio.cpp:1052-1108: greedy batches over the length-sorted queue.
    Returns a list of batches (lists of queue positions)."""
    max_acc = 300 * 1000 * seg_sample
    max_sent = 60 * 1000 * seg_sample
    q = list(range(len(lengths_sorted)))
    out = []
    while q:
        bs_acc = 0
        max_len = 0
        max_batch = batch_size if use_gpu else 1
        max_batch = min(max_batch, len(q))
        batch = []
        for _ in range(max_batch):
            length = lengths_sorted[q[0]]
            if length >= max_sent:
                if bs_acc == 0:
                    bs_acc += 1
                    batch.append(q.pop(0))
                break
            max_len = max(max_len, length)
            if max_len * (bs_acc + 1) > max_acc:
                break
            bs_acc += 1
            batch.append(q.pop(0))
        if not batch:
            break
        out.append(batch)
    return out
