"""Offline BERT tokenizer: basic (whitespace / punctuation / accent / CJK) + greedy WordPiece.

Capability parity with the vendored scaelum/dataset/glue/tokenization.py:40-408 minus every
network path (no ``from_pretrained`` URL map: a local ``vocab.txt`` is required).
"""
from __future__ import annotations

import collections
import unicodedata
from typing import Dict, List


def load_vocab(vocab_file: str) -> "collections.OrderedDict[str, int]":
    vocab: "collections.OrderedDict[str, int]" = collections.OrderedDict()
    with open(vocab_file, "r", encoding="utf-8") as reader:
        for index, line in enumerate(reader):
            token = line.rstrip("\n")
            if token == "" and index > 0:
                continue
            vocab[token.strip()] = index
    return vocab


def whitespace_tokenize(text: str) -> List[str]:
    text = text.strip()
    return text.split() if text else []


def _is_whitespace(ch: str) -> bool:
    return ch in (" ", "\t", "\n", "\r") or unicodedata.category(ch) == "Zs"


def _is_control(ch: str) -> bool:
    if ch in ("\t", "\n", "\r"):
        return False
    return unicodedata.category(ch).startswith("C")


def _is_punctuation(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF
            or 0x2A700 <= cp <= 0x2B73F or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF
            or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


class BasicTokenizer:
    def __init__(self, do_lower_case: bool = True,
                 never_split=("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]")):
        self.do_lower_case = do_lower_case
        self.never_split = set(never_split)

    def tokenize(self, text: str) -> List[str]:
        cleaned = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            if _is_cjk(cp):
                cleaned.append(f" {ch} ")
            elif _is_whitespace(ch):
                cleaned.append(" ")
            else:
                cleaned.append(ch)
        out: List[str] = []
        for token in whitespace_tokenize("".join(cleaned)):
            if token in self.never_split:
                out.append(token)
                continue
            if self.do_lower_case:
                token = "".join(c for c in unicodedata.normalize("NFD", token.lower())
                                if unicodedata.category(c) != "Mn")
            word = ""
            for ch in token:
                if _is_punctuation(ch):
                    if word:
                        out.append(word)
                        word = ""
                    out.append(ch)
                else:
                    word += ch
            if word:
                out.append(word)
        return out


class WordpieceTokenizer:
    def __init__(self, vocab: Dict[str, int], unk_token: str = "[UNK]",
                 max_input_chars_per_word: int = 100):
        self.vocab = vocab
        self.unk_token = unk_token
        self.max_input_chars_per_word = max_input_chars_per_word

    def tokenize(self, text: str) -> List[str]:
        output: List[str] = []
        for token in whitespace_tokenize(text):
            if len(token) > self.max_input_chars_per_word:
                output.append(self.unk_token)
                continue
            pieces, start, bad = [], 0, False
            while start < len(token):
                end, cur = len(token), None
                while start < end:
                    sub = token[start:end]
                    if start > 0:
                        sub = "##" + sub
                    if sub in self.vocab:
                        cur = sub
                        break
                    end -= 1
                if cur is None:
                    bad = True
                    break
                pieces.append(cur)
                start = end
            output.extend([self.unk_token] if bad else pieces)
        return output


class BertTokenizer:
    def __init__(self, vocab_file: str, do_lower_case: bool = True, max_len: int = 512,
                 never_split=("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]")):
        self.vocab = load_vocab(vocab_file)
        self.ids_to_tokens = collections.OrderedDict((i, t) for t, i in self.vocab.items())
        self.basic_tokenizer = BasicTokenizer(do_lower_case=do_lower_case, never_split=never_split)
        self.wordpiece_tokenizer = WordpieceTokenizer(vocab=self.vocab)
        self.max_len = max_len

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for token in self.basic_tokenizer.tokenize(text):
            out.extend(self.wordpiece_tokenizer.tokenize(token))
        return out

    def convert_tokens_to_ids(self, tokens: List[str]) -> List[int]:
        ids = [self.vocab.get(t, self.vocab.get("[UNK]", 0)) for t in tokens]
        if len(ids) > self.max_len:
            raise ValueError(f"sequence length {len(ids)} exceeds the model maximum {self.max_len}")
        return ids

    def convert_ids_to_tokens(self, ids: List[int]) -> List[str]:
        return [self.ids_to_tokens[i] for i in ids]
