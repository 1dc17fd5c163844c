"""Text front end of the hot path: cleaning, number verbalisation and character tokenisation for
Mongolian (Khalkha) and Kazakh Cyrillic. Host-side, pure Python; the ids it produces must be
bit-identical to the reference's (north_star), which tests/test_text_golden.py checks against
vectors generated from the live reference.

Behavioural contract restated from the reference:
  * vocabulary / encode        — src/utils/tokenizer.py:8-13, 32-55, 80-106
  * cleaning pipeline order    — src/utils/text_cleaner.py:120-130 (NFC, punctuation map, abbreviations,
                                 numbers, character filter, whitespace, repeated punctuation, lower-case)
  * number verbalisation       — src/utils/number_norm.py:239-348 (cardinal / attributive / ordinal forms)
  * ordered rewrite passes     — src/utils/number_norm.py:382-566

The implementation is table-driven: each language is a ``Lexicon`` record and the rewrite passes are
an ordered list of (compiled regex, handler) pairs.
"""

from __future__ import annotations

import re
import unicodedata
from dataclasses import dataclass, field

LANGS = frozenset({"mn", "kz"})


def validate_language(lang: str) -> str:
    if lang not in LANGS:
        raise ValueError(f"Unsupported language '{lang}'. Expected one of: {', '.join(sorted(LANGS))}")
    return lang


# ------------------------------------------------------------------------------------------------
# tokenizer
# ------------------------------------------------------------------------------------------------
SPECIAL_TOKENS = ["<PAD>", "<BOS>", "<EOS>", "<UNK>", "[LANG_MN]", "[LANG_KZ]", "[FEMALE]", "[MALE]", "[YOUNG]",
                  "[MIDDLE]", "[ELDERLY]"]
_LETTERS_MN = "абвгдеёжзийклмноөпрстуүфхцчшщъыьэюя"
_LETTERS_KZ_ONLY = "әғқңұһі"
_PUNCT = " .,!?-:;\"'()"
VOCAB = SPECIAL_TOKENS + list(_LETTERS_MN + _LETTERS_KZ_ONLY + _PUNCT)


class CyrillicTokenizer:
    """Character-level tokenizer: ``[LANG] [attr...] chars`` with <UNK> (3) for anything else."""

    def __init__(self) -> None:
        self._to_id = {tok: i for i, tok in enumerate(VOCAB)}
        self._to_tok = dict(enumerate(VOCAB))
        self.pad_id, self.bos_id, self.eos_id, self.unk_id = (self._to_id[t] for t in SPECIAL_TOKENS[:4])

    @property
    def vocab_size(self) -> int:
        return len(VOCAB)

    def encode(self, text: str, lang: str = "mn", attr_tokens: list[str] | None = None) -> list[int]:
        lang = validate_language(lang)
        out = [self._to_id["[LANG_MN]" if lang == "mn" else "[LANG_KZ]"]]
        out.extend(self._to_id.get(a, self.unk_id) for a in (attr_tokens or ()))
        out.extend(self._to_id.get(ch, self.unk_id) for ch in text)
        return out

    def decode(self, ids: list[int]) -> str:
        toks = (self._to_tok.get(i, "<UNK>") for i in ids)
        return "".join(t for t in toks if t not in SPECIAL_TOKENS)

    def token_to_id(self, token: str) -> int:
        return self._to_id.get(token, self.unk_id)

    def id_to_token(self, idx: int) -> str:
        return self._to_tok.get(idx, "<UNK>")


# ------------------------------------------------------------------------------------------------
# number verbalisation
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Lexicon:
    """Per-language word tables. Each entry is (standalone form, attributive form)."""

    code: str
    ones: tuple
    ten: tuple
    tens: dict
    hundred: tuple
    scales: tuple  # ((value, (standalone, attributive)), ...) descending
    ordinal_by_vowel: dict
    ordinal_default: str
    zero: str
    minus: str
    point: str
    percent: str
    year: str
    month: str
    hour: str
    minute: str = "минут"
    second: str = "секунд"
    degree: str = "градус"
    half: str = ""
    range_from: str = ""
    range_to: str = ""
    column: int = 0  # which column of the shared symbol tables applies
    extra: dict = field(default_factory=dict)


def _pairs(s: str) -> tuple:
    return tuple(tuple(p.split("/")) if "/" in p else (p, p) for p in s.split())


MN = Lexicon(
    code="mn",
    ones=(("", ""),) + _pairs("нэг хоёр гурав/гурван дөрөв/дөрвөн тав/таван зургаа/зургаан долоо/долоон найм/найман ес/есөн"),
    ten=("арав", "арван"),
    tens=dict(zip(range(2, 10), _pairs("хорь/хорин гуч/гучин дөч/дөчин тавь/тавин жар/жаран дал/далан ная/наян ер/ерэн"))),
    hundred=("зуу", "зуун"),
    scales=((10**12, ("их наяд", "их наяд")), (10**9, ("тэрбум", "тэрбум")), (10**6, ("сая", "сая")),
            (10**3, ("мянга", "мянган"))),
    ordinal_by_vowel={**dict.fromkeys("аоуь", "дугаар"), **dict.fromkeys("эөүие", "дүгээр")},
    ordinal_default="дугаар",
    zero="тэг", minus="хасах", point="цэг", percent="хувь", year="оны", month="сарын", hour="цаг",
    half="хагас", range_from="аас", range_to="хүртэл", column=0,
)
KZ = Lexicon(
    code="kz",
    ones=(("", ""),) + _pairs("бір екі үш төрт бес алты жеті сегіз тоғыз"),
    ten=("он", "он"),
    tens=dict(zip(range(2, 10), _pairs("жиырма отыз қырық елу алпыс жетпіс сексен тоқсан"))),
    hundred=("жүз", "жүз"),
    scales=((10**9, ("миллиард", "миллиард")), (10**6, ("миллион", "миллион")), (10**3, ("мың", "мың"))),
    ordinal_by_vowel=dict.fromkeys("аеыіоөұү", "нші"),
    ordinal_default="нші",
    zero="нөл", minus="минус", point="бүтін", percent="пайыз", year="жылдың", month="айдың", hour="сағат",
    half="жарты", range_from="ден", range_to="дейін", column=1,
)
LEXICONS = {"mn": MN, "kz": KZ}

CURRENCY_SIGNS = {"₮": "төгрөг", "₸": "теңге", "$": "доллар", "€": "евро", "£": "фунт", "¥": "иен", "₽": "рубль"}
CURRENCY_ISO = {"MNT": "төгрөг", "KZT": "теңге", "USD": "доллар", "EUR": "евро", "GBP": "фунт", "JPY": "иен",
                "CNY": "юань", "RUB": "рубль", "KRW": "вон"}
# symbol -> (mn, kz); insertion order is the replacement order
MATH_WORDS = {
    "+": ("нэмэх", "қосу"), "×": ("үржүүлэх", "көбейту"), "÷": ("хуваах", "бөлу"), "=": ("тэнцүү", "тең"),
    "≠": ("тэнцүү биш", "тең емес"), "<": ("бага", "кіші"), ">": ("их", "үлкен"),
    "≤": ("бага буюу тэнцүү", "кіші немесе тең"), "≥": ("их буюу тэнцүү", "үлкен немесе тең"),
    "±": ("нэмэх хасах", "плюс минус"), "~": ("ойролцоогоор", "шамамен"),
}
_ROMAN_TOKEN = re.compile(r"\b(M{0,3}(?:CM|CD|D?C{0,3})(?:XC|XL|L?X{0,3})(?:IX|IV|V?I{0,3}))\b")
_ROMAN_DIGITS = (("M", 1000), ("CM", 900), ("D", 500), ("CD", 400), ("C", 100), ("XC", 90), ("L", 50), ("XL", 40),
                 ("X", 10), ("IX", 9), ("V", 5), ("IV", 4), ("I", 1))


def roman_value(s: str) -> int | None:
    if not s:
        return None
    total, pos = 0, 0
    for sym, val in _ROMAN_DIGITS:
        while s.startswith(sym, pos):
            total += val
            pos += len(sym)
    return total if (pos == len(s) and total > 0) else None


class NumberNormalizer:
    """Spells integers (cardinal / attributive / ordinal) and rewrites numeric expressions in text."""

    def __init__(self, lang: str = "mn") -> None:
        self._lang = validate_language(lang)
        self._lx = LEXICONS[self._lang]
        self._rules = self._build_rules()

    @property
    def lang(self) -> str:
        return self._lang

    @lang.setter
    def lang(self, value: str) -> None:
        value = validate_language(value)
        if value != self._lang:
            self._lang, self._lx = value, LEXICONS[value]
            self._rules = self._build_rules()

    # ---- integer -> words ------------------------------------------------------------------------
    def _below_100(self, n: int, col: int) -> str:
        lx = self._lx
        if n == 0:
            return ""
        if n < 10:
            return lx.ones[n][col]
        if n == 10:
            return lx.ten[col]
        tens, ones = divmod(n, 10)
        if tens == 1:
            return f"{lx.ten[1]} {lx.ones[ones][col]}"
        if ones == 0:
            return lx.tens[tens][col]
        return f"{lx.tens[tens][1]} {lx.ones[ones][col]}"

    def _below_1000(self, n: int, col: int) -> str:
        lx = self._lx
        if n < 100:
            return self._below_100(n, col)
        h, rest = divmod(n, 100)
        lead = "" if h == 1 else lx.ones[h][1] + " "
        if rest == 0:
            return lead + lx.hundred[col]
        return f"{lead}{lx.hundred[1]} {self._below_100(rest, col)}"

    def _spell(self, n: int, col: int) -> str:
        if n < 1000:
            return self._below_1000(n, col)
        words: list[str] = []
        left = n
        for value, forms in self._lx.scales:
            if left < value:
                continue
            count, left = divmod(left, value)
            form = forms[1] if (col == 1 and left == 0) else forms[0]
            words.append(form if count == 1 else f"{self._spell(count, 1)} {form}")
        if left > 0:
            words.append(self._below_1000(left, col))
        return " ".join(words)

    def _signed(self, n: int, col: int) -> str:
        if n == 0:
            return self._lx.zero
        if n < 0:
            return f"{self._lx.minus} {self._signed(-n, col)}"
        return self._spell(n, col)

    def convert(self, n: int) -> str:
        return self._signed(n, 0)

    def convert_attributive(self, n: int) -> str:
        return self._signed(n, 1)

    def convert_ordinal(self, n: int) -> str:
        word = self.convert(n)
        suffix = next((self._lx.ordinal_by_vowel[c] for c in reversed(word.lower()) if c in self._lx.ordinal_by_vowel),
                      self._lx.ordinal_default)
        return word + suffix

    # ---- rewrite passes ---------------------------------------------------------------------------
    def _spell_digits(self, digits: str) -> str:
        return " ".join(self._lx.zero if d == "0" else self.convert(int(d)) for d in digits)

    def _money(self, token: str) -> str:
        if token in CURRENCY_SIGNS:
            return CURRENCY_SIGNS[token]
        return CURRENCY_ISO.get(token.upper(), token)

    def _build_rules(self) -> list:
        lx = self._lx
        card, attr, ordn = self.convert, self.convert_attributive, self.convert_ordinal
        signs = "|".join(re.escape(s) for s in CURRENCY_SIGNS)
        codes = "|".join(CURRENCY_ISO)

        def date(y: str, mo: str, d: str) -> str:
            return f"{attr(int(y))} {lx.year} {ordn(int(mo))} {lx.month} {card(int(d))}"

        def clock(m: re.Match) -> str:
            parts = [f"{attr(int(m[1]))} {lx.hour}", f"{attr(int(m[2]))} {lx.minute}"]
            if m[3] is not None:
                parts.append(f"{attr(int(m[3]))} {lx.second}")
            return " ".join(parts)

        def temperature(m: re.Match) -> str:
            parts = [lx.minus] if m[1] == "-" else []
            parts.append(f"{attr(int(m[2]))} {lx.degree}")
            unit = (m[3] or "").upper()
            if unit == "C":
                parts.append("цельсий")
            elif unit == "F":
                parts.append("фаренгейт")
            return " ".join(parts)

        def fraction(m: re.Match) -> str:
            num, den = int(m[1]), int(m[2])
            if (num, den) == (1, 2):
                return lx.half
            if lx.code == "mn":
                o = ordn(den)
                return f"{o}{'ийн' if o.endswith('дүгээр') else 'ын'} {card(num)}"
            return f"{card(den)} ден {card(num)}"

        def phone(m: re.Match) -> str:
            return f"{MATH_WORDS['+'][lx.column]} " + self._spell_digits(re.sub(r"\D", "", m[0][1:]))

        def roman(m: re.Match) -> str:
            v = roman_value(m[1])
            return m[0] if v is None else ordn(v)

        def math_symbols(text: str) -> str:
            for sym, words in MATH_WORDS.items():
                if sym in text:
                    text = text.replace(sym, f" {words[lx.column]} ")
            return text

        R = re.compile
        return [
            # thousands separators: 1,234,567 / 1 234 567
            (R(r"(\d{1,3})(?:[ ,](\d{3}))+"), lambda m: m[0].replace(",", "").replace(" ", "")),
            # dates: Y-M-D then D-M-Y, separators / . -
            (R(r"(\d{4})[/.\-](\d{1,2})[/.\-](\d{1,2})"), lambda m: date(m[1], m[2], m[3])),
            (R(r"(\d{1,2})[/.\-](\d{1,2})[/.\-](\d{4})"), lambda m: date(m[3], m[2], m[1])),
            (R(r"(\d{1,2}):(\d{2})(?::(\d{2}))?"), clock),
            (R(r"(-?)(\d+)°\s*([CcFf])?"), temperature),
            # currency: amount then sign / ISO code, then sign before amount
            (R(rf"(\d+)\s*({signs}|(?:{codes})(?!\w))"), lambda m: f"{attr(int(m[1]))} {self._money(m[2])}"),
            (R(rf"({signs})\s*(\d+)"), lambda m: f"{attr(int(m[2]))} {self._money(m[1])}"),
            (R(r"(\d+)%"), lambda m: f"{attr(int(m[1]))} {lx.percent}"),
            (R(r"(\d+)\.(\d+)"), lambda m: f"{card(int(m[1]))} {lx.point} " + " ".join(card(int(d)) for d in m[2])),
            (R(r"(\d{1,2})/(\d{1,2})"), fraction),
            (R(r"\+\d[\d\s\-]{6,15}\d"), phone),
            (R(r"(\d+)\s*[-–—]\s*(\d+)"), lambda m: f"{card(int(m[1]))} {lx.range_from} {card(int(m[2]))} {lx.range_to}"),
            # written ordinals: 20-р, 3-дугаар, 5-ші
            (R(r"(\d+)-р\b"), lambda m: ordn(int(m[1]))),
            (R(r"(\d+)-д(?:угаар|үгээр|ахь)"), lambda m: ordn(int(m[1]))),
            (R(r"(\d+)-(?:ші|шы)"), lambda m: ordn(int(m[1]))),
            # genitive markers read as the attributive cardinal
            (R(r"(\d+)-(?:ны|ний|ын|ийн)\b"), lambda m: attr(int(m[1]))),
            (_ROMAN_TOKEN, roman),
            math_symbols,
            # a number directly before a (lower-case) Cyrillic word is attributive; everything left is cardinal
            (R(r"(\d+)(?=\s+[а-яёәғқңұһі])"), lambda m: attr(int(m[1]))),
            (R(r"\d+"), lambda m: card(int(m[0]))),
        ]

    def normalize_text(self, text: str) -> str:
        for rule in self._rules:
            text = rule(text) if callable(rule) else rule[0].sub(rule[1], text)
        return text


# ------------------------------------------------------------------------------------------------
# cleaner
# ------------------------------------------------------------------------------------------------
_PUNCT_FOLD = {"…": "...", "–": "-", "—": "-", "«": '"', "»": '"', "“": '"', "”": '"', "‘": "'",
               "„": '"'}
_KEEP = frozenset(_LETTERS_MN + _LETTERS_MN.upper() + _LETTERS_KZ_ONLY + _LETTERS_KZ_ONLY.upper() + _PUNCT)
_METRIC = {"км": "километр", "см": "сантиметр", "кг": "килограмм", "мл": "миллилитр"}
# insertion order matters (sequential substitutions)
_ABBREV = {
    "mn": {"г.": "оны", **_METRIC, "т.": "товч", "тов.": "товч", "ж.": "жил", "сар.": "сар", "өд.": "өдөр",
           "мин.": "минут", "сек.": "секунд", "цаг.": "цаг"},
    "kz": {"ж.": "жыл", **_METRIC, "мин.": "минут", "сек.": "секунд", "сағ.": "сағат"},
}
_UNIT_AFTER_DIGIT = {"м": "метр", "г": "грамм", "л": "литр"}


class TextCleaner:
    def __init__(self) -> None:
        self._numbers = {code: NumberNormalizer(code) for code in ("mn", "kz")}
        self._tokenizer = CyrillicTokenizer()
        self._abbrev_rules = {
            code: [(re.compile(rf"(?<!\w){re.escape(k)}(?!\w)", re.IGNORECASE), v) for k, v in table.items()]
            for code, table in _ABBREV.items()
        }
        self._unit_rules = [(re.compile(rf"(\d)\s*{re.escape(k)}(?!\w)", re.IGNORECASE), rf"\1 {v}")
                            for k, v in _UNIT_AFTER_DIGIT.items()]
        self._ws = re.compile(r"\s+")
        self._repeat = re.compile(r"([.!?,]){2,}")

    def normalize_unicode(self, text: str) -> str:
        return unicodedata.normalize("NFC", text)

    def replace_punctuation(self, text: str) -> str:
        for src, dst in _PUNCT_FOLD.items():
            text = text.replace(src, dst)
        return text

    def expand_abbreviations(self, text: str, lang: str = "mn") -> str:
        lang = validate_language(lang)
        for rx, full in self._abbrev_rules[lang]:
            text = rx.sub(full, text)
        for rx, repl in self._unit_rules:
            text = rx.sub(repl, text)
        return text

    def remove_invalid_chars(self, text: str) -> str:
        return "".join(filter(_KEEP.__contains__, text))

    def normalize_whitespace(self, text: str) -> str:
        return self._ws.sub(" ", text).strip()

    def normalize_punctuation(self, text: str) -> str:
        return self._repeat.sub(r"\1", text)

    def clean(self, text: str, lang: str = "mn") -> str:
        lang = validate_language(lang)
        text = self.replace_punctuation(self.normalize_unicode(text))
        text = self.expand_abbreviations(text, lang=lang)
        text = self._numbers[lang].normalize_text(text)
        text = self.normalize_punctuation(self.normalize_whitespace(self.remove_invalid_chars(text)))
        return text.lower()

    def text_to_sequence(self, text: str, lang: str = "mn", attr_tokens: list[str] | None = None) -> list[int]:
        return self._tokenizer.encode(self.clean(text, lang=lang), lang=lang, attr_tokens=attr_tokens)

    @property
    def vocab_size(self) -> int:
        return self._tokenizer.vocab_size
