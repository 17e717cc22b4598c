"""TEST INFRASTRUCTURE — golden vectors for the WordPiece tokenizer / bertify drop-in, produced by the
UNMODIFIED reference (Models/Bert/tokenization.py, Utils/VQA_Dataset.py:415-436).  Build container only:

    python -m oracle.gen_tokenizer_golden

Writes tests/golden/tokenizer_vocab.txt (a small synthetic vocabulary) and
tests/golden/tokenizer_golden.json (strings -> tokens, word lists -> (ids, offsets)).
"""
import json
import os
import random
import sys
import types

from . import ref_harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SPECIAL = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
WORDS = ["the", "of", "and", "stop", "coca", "cola", "##ing", "##ed", "##s", "un", "##aff", "##able", "exit", "open",
         "2019", "##19", "!", ".", ",", "-", "'", "(", ")", "中", "国", "cafe", "##cafe", "shop", "##shop",
         "street", "##street", "<", ">", "ocr", "od", "q", "ß", "##ß", "no", "##vember", "sale", "##sale"]
ALNUM = "abcdefghijklmnopqrstuvwxyz0123456789"


def vocabulary():
    return SPECIAL + list(ALNUM) + ["##" + c for c in ALNUM] + WORDS


def random_strings(n, seed=7):
    rng = random.Random(seed)
    alphabet = list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789") * 3 + \
        list(" \t\n.,!-'\"()[]{}^$`~<>") * 2 + \
        list("éèüñçÅøß中国日本語 ​́�\x00\x07 İǅ΅")
    out = ["", " ", "Stop!", "Coca-Cola", "unaffable", "UNAFFABLE shops", "café street", "中国2019", "<OCR>", "<OD>",
           "a" * 100, "a" * 101, "No.19 Street's SALE!!", "opened", "éxit"]
    for _ in range(n):
        out.append("".join(rng.choice(alphabet) for _ in range(rng.randint(0, 40))))
    return out


def main():
    if not ref_harness.available():
        raise SystemExit("needs the reference tree (build container only)")
    ref_harness._install_stubs()
    if ref_harness.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_harness.REF_ROOT)
    from Models.Bert import tokenization as ref_tok
    from Utils.VQA_Dataset import VQA_Dataset
    vocab_file = os.path.join(OUT, "tokenizer_vocab.txt")
    with open(vocab_file, "w", encoding="utf8") as f:
        f.write("\n".join(vocabulary()) + "\n")
    tok = ref_tok.BertTokenizer(vocab_file)
    strings = random_strings(400)
    tokens = [tok.tokenize(s) for s in strings]
    rng = random.Random(11)
    items = [[rng.choice(strings[2:60]) for _ in range(rng.randint(0, 4))] for _ in range(60)]
    fake = types.SimpleNamespace(bert_tokenizer=tok)
    bert = [VQA_Dataset.bertify(fake, list(words)) for words in items]
    bert_str = [VQA_Dataset.bertify(fake, s) for s in strings[2:12]]
    with open(os.path.join(OUT, "tokenizer_golden.json"), "w", encoding="utf8") as f:
        json.dump({"strings": strings, "tokens": tokens, "items": items,
                   "bertify": [[list(i), o] for i, o in bert],
                   "bertify_str": [[list(i), o] for i, o in bert_str]}, f, ensure_ascii=True)
    print("wrote", len(strings), "strings,", len(items), "items")


if __name__ == "__main__":
    main()
