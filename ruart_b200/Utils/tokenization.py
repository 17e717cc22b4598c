"""Drop-in for the reference's Models/Bert/tokenization.py (`BertTokenizer`, `BasicTokenizer`,
`WordpieceTokenizer`, `load_vocab`, `whitespace_tokenize`) and for `VQA_Dataset.bertify`
(Utils/VQA_Dataset.py:415-436) — the host string work immediately before the hot path
(SURVEY.md §8f-4).  Same results, different mechanics:

  * characters are classified once (drop / space / punctuation / CJK / keep) through a memoised
    table instead of three `unicodedata` calls per character per pass;
  * the basic tokenizer is a single pass over the text that cleans, isolates CJK characters and
    punctuation, lower-cases and strips accents (Models/Bert/tokenization.py:164-264);
  * WordPiece is the reference's greedy longest-match-first rule (tokenization.py:274-325) with the
    search bounded by the longest vocabulary entry, and every distinct word is tokenized once
    (OCR tokens and object labels repeat heavily within a batch);
  * `bertify_batch` tokenizes all items of a list at once and returns what the collate step needs:
    padded ids + mask, the per-word `[st, ed)` offsets as nested lists (the reference's layout) and
    as the CSR array of ruart_b200/Utils/collate.py.
"""
import collections
import os
import unicodedata

import numpy as np

_DROP, _SPACE, _PUNCT, _CJK, _KEEP = range(5)

_CJK_RANGES = ((0x4E00, 0x9FFF), (0x3400, 0x4DBF), (0x20000, 0x2A6DF), (0x2A700, 0x2B73F), (0x2B740, 0x2B81F),
               (0x2B820, 0x2CEAF), (0xF900, 0xFAFF), (0x2F800, 0x2FA1F))


def convert_to_unicode(text):
    if isinstance(text, str):
        return text
    if isinstance(text, bytes):
        return text.decode("utf-8", "ignore")
    raise ValueError("Unsupported string type: %s" % (type(text)))


printable_text = convert_to_unicode


def load_vocab(vocab_file):
    """token -> index, one token per line (tokenization.py:62-74)."""
    vocab = collections.OrderedDict()
    with open(vocab_file, "r", encoding="utf8") as reader:
        for index, line in enumerate(reader):
            vocab[line.strip()] = index
    return vocab


def whitespace_tokenize(text):
    return text.split()


class _CharClass(dict):
    """char -> class, computed on first use (tokenization.py:252-264, 217-250, 328-362)."""

    def __missing__(self, ch):
        cp = ord(ch)
        if ch in " \t\n\r":
            c = _SPACE
        elif cp == 0 or cp == 0xFFFD:
            c = _DROP
        else:
            cat = unicodedata.category(ch)
            if cat.startswith("C"):
                c = _DROP
            elif cat == "Zs":
                c = _SPACE
            elif any(lo <= cp <= hi for lo, hi in _CJK_RANGES):
                c = _CJK
            elif (33 <= cp <= 47) or (58 <= cp <= 64) or (91 <= cp <= 96) or (123 <= cp <= 126) or cat.startswith("P"):
                c = _PUNCT
            else:
                c = _KEEP
        self[ch] = c
        return c


_CLASS = _CharClass()


def _is_whitespace(char):
    return _CLASS[char] == _SPACE


def _is_control(char):
    return char not in "\t\n\r" and unicodedata.category(char).startswith("C")


def _is_punctuation(char):
    cp = ord(char)
    if (33 <= cp <= 47) or (58 <= cp <= 64) or (91 <= cp <= 96) or (123 <= cp <= 126):
        return True
    return unicodedata.category(char).startswith("P")


class BasicTokenizer(object):
    """Punctuation splitting, lower casing, accent stripping (tokenization.py:153-264)."""

    def __init__(self, do_lower_case=True):
        self.do_lower_case = do_lower_case

    def _fold(self, word):
        """lower() + NFD + drop combining marks; may expose new punctuation / spaces, so the result is
        re-split by the caller exactly like the reference's order of operations."""
        word = word.lower()
        if word.isascii():
            return word
        return "".join(ch for ch in unicodedata.normalize("NFD", word) if unicodedata.category(ch) != "Mn")

    def tokenize(self, text):
        text = convert_to_unicode(text)
        cls = _CLASS
        # pass 1 (clean + CJK isolation + whitespace split): words of KEEP / PUNCT characters
        words, cur = [], []
        for ch in text:
            c = cls[ch]
            if c == _DROP:
                continue
            if c == _SPACE:
                if cur:
                    words.append("".join(cur))
                    cur = []
            elif c == _CJK:
                if cur:
                    words.append("".join(cur))
                    cur = []
                words.append(ch)
            else:
                cur.append(ch)
        if cur:
            words.append("".join(cur))
        # pass 2 (per word: fold, then split on punctuation; a fold can produce whitespace, which the
        # reference removes with its final whitespace_tokenize)
        out = []
        for w in words:
            if self.do_lower_case:
                w = self._fold(w)
            piece = []
            for ch in w:
                if _is_punctuation(ch):
                    if piece:
                        out.append("".join(piece))
                        piece = []
                    out.append(ch)
                else:
                    piece.append(ch)
            if piece:
                out.append("".join(piece))
        # the reference re-splits on str.split() whitespace at the end (tokenization.py:183): characters
        # such as U+2028 are whitespace to str.split() but not to _is_whitespace, and a fold can expose some
        if any(ch.isspace() for t in out for ch in t):
            out = " ".join(out).split()
        return out


class WordpieceTokenizer(object):
    """Greedy longest-match-first WordPiece (tokenization.py:266-325)."""

    def __init__(self, vocab, unk_token="[UNK]", max_input_chars_per_word=100):
        self.vocab = vocab
        self.unk_token = unk_token
        self.max_input_chars_per_word = max_input_chars_per_word
        self._longest = max((len(t) for t in vocab), default=0)
        self._cache = {}

    def _word(self, token):
        hit = self._cache.get(token)
        if hit is not None:
            return hit
        n = len(token)
        if n > self.max_input_chars_per_word:
            res = (self.unk_token,)
        else:
            vocab, longest = self.vocab, self._longest
            pieces, start = [], 0
            while start < n:
                # no vocabulary entry is longer than `longest` characters ("##" included)
                end = min(n, start + longest)
                found = None
                while end > start:
                    cand = token[start:end] if start == 0 else "##" + token[start:end]
                    if cand in vocab:
                        found = cand
                        break
                    end -= 1
                if found is None:
                    pieces = None
                    break
                pieces.append(found)
                start = end
            res = (self.unk_token,) if pieces is None else tuple(pieces)
        if len(self._cache) < 1 << 20:
            self._cache[token] = res
        return res

    def tokenize(self, text):
        out = []
        for token in convert_to_unicode(text).split():
            out.extend(self._word(token))
        return out


class BertTokenizer(object):
    """End-to-end tokenization (tokenization.py:86-151)."""

    def __init__(self, vocab_file, do_lower_case=True):
        if not os.path.isfile(vocab_file):
            raise ValueError("Can't find a vocabulary file at path '{}'.".format(vocab_file))
        self.vocab = load_vocab(vocab_file)
        self.ids_to_tokens = collections.OrderedDict((i, t) for t, i in self.vocab.items())
        self.basic_tokenizer = BasicTokenizer(do_lower_case=do_lower_case)
        self.wordpiece_tokenizer = WordpieceTokenizer(vocab=self.vocab)
        self._text_cache = {}

    def tokenize(self, text):
        hit = self._text_cache.get(text)
        if hit is None:
            hit = []
            for token in self.basic_tokenizer.tokenize(text):
                hit.extend(self.wordpiece_tokenizer._word(token))
            if len(self._text_cache) < 1 << 20:
                self._text_cache[text] = hit
        return list(hit)

    def convert_tokens_to_ids(self, tokens):
        vocab = self.vocab
        return [vocab[t] for t in tokens]

    def convert_ids_to_tokens(self, ids):
        table = self.ids_to_tokens
        return [table[i] for i in ids]

    @classmethod
    def from_pretrained(cls, pretrained_model_name, do_lower_case=True):
        """A local vocabulary file (the reference also resolves hub names to URLs; there is no network)."""
        try:
            return cls(pretrained_model_name, do_lower_case)
        except (FileNotFoundError, ValueError):
            return None

    # ------------------------------------------------------------------ batch form
    def bertify(self, words):
        """VQA_Dataset.bertify (VQA_Dataset.py:415-436): [CLS] pieces... [SEP] ids and the [st, ed)
        piece range of every word; a plain string gives ids only (empty offsets)."""
        bpe, offsets = ["[CLS]"], []
        if isinstance(words, list):
            for word in words:
                now = self.tokenize(word)
                offsets.append([len(bpe), len(bpe) + len(now)])
                bpe.extend(now)
            if len(words) == 0:
                offsets = [1, 1]
        elif isinstance(words, str):
            bpe = bpe + self.tokenize(words)
        else:
            raise AssertionError("BERT tokenizer is wrong")
        bpe.append("[SEP]")
        return self.convert_tokens_to_ids(bpe), offsets

    def bertify_batch(self, items, max_len=None):
        """All items of a list at once: `items` = list of word lists.  Returns a dict with
        `bert` int64 [N, L] (0-padded, L = max_len or the longest item), `bert_mask` bool [N, L],
        `bert_offsets` (the reference's nested lists) and `bert_offsets_csr` int32 [4, n_words]
        (item, word, st, ed) — the layout of VQA_collate_fun (VQA_Dataset.py:505-517) for these keys."""
        ids_all, offs_all = [], []
        for words in items:
            ids, offs = self.bertify(list(words))
            ids_all.append(ids)
            offs_all.append(offs)
        longest = max((len(i) for i in ids_all), default=0)
        L = longest if max_len is None else max_len
        if longest > L:
            raise ValueError("an item has %d wordpieces, more than max_len=%d" % (longest, L))
        bert = np.zeros((len(items), L), dtype=np.int64)
        for r, ids in enumerate(ids_all):
            bert[r, :len(ids)] = ids
        rows = [(r, j, st, ed) for r, offs in enumerate(offs_all) if offs and isinstance(offs[0], list)
                for j, (st, ed) in enumerate(offs)]
        rows += [(r, 0, 1, 1) for r, offs in enumerate(offs_all) if offs == [1, 1]]
        rows.sort()
        csr = np.asarray(rows, dtype=np.int32).reshape(-1, 4).T.copy()
        return {"bert": bert, "bert_mask": bert != 0, "bert_offsets": offs_all, "bert_offsets_csr": csr}
