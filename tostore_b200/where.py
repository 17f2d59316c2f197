"""Structured WHERE prefilter: condition tree -> postfix program for
`tsc_index_filter_where[_text]` (the GPU evaluates it over attribute columns, tsc_where.cuh).

Host mirror of the reference's condition handling for integer / double / text fields (paths
relative to /root/reference/lib/src):
  * `QueryCondition` map form (`{'AND': [...]}`, `{'OR': [...]}`, leaf
    `{field: {op: value}}` / `{field: value}`)      query/query_condition.dart:24-55, :486-520
  * operand normalisation to the field's type       query/query_condition.dart:743-815,
                                                    model/table_schema.dart:1356-1421
  * operator meaning                                handler/value_matcher.dart:570-612
An operator map with several entries is an OR of them (value_matcher.dart:552-563);
several fields in one leaf are an AND (:499-510).

Text fields: operands go through `convertValue` for DataType.text — `toString().trim()`
(model/table_schema.dart:1421-1442; this includes LIKE patterns, which the reference trims
like any other operand) — and travel as UTF-16 code units, a Dart String's own form, so the
GPU's comparisons are `String.compareTo` (value_matcher.dart:211-240) and LIKE is
`ValueMatcher.matchesLike` (:318-331).

The reference has no WHERE for `vectorSearch`; this is the additive prefilter of
BASELINE config 5.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Tuple

import numpy as np

COL_I64, COL_F64, COL_TEXT = 0, 1, 2
COL_BOOL = 16     # host-side only: a boolean field is an int64 column of 0 / 1 on the device
COL_DATETIME = 17  # host-side only: a datetime field is a text column of ISO-8601 strings
W_LEAF, W_AND, W_OR = 0, 1, 2
(OP_EQ, OP_NE, OP_GT, OP_GE, OP_LT, OP_LE, OP_BETWEEN, OP_IN, OP_NOT_IN, OP_IS_NULL,
 OP_IS_NOT_NULL, OP_TRUE, OP_FALSE, OP_LIKE, OP_NOT_LIKE) = range(15)
MAX_OPS, MAX_IN_ARGS, MAX_TEXTS = 64, 4096, 4096
_INT64_MIN, _INT64_MAX = -(1 << 63), (1 << 63) - 1
_SIMPLE = {"=": OP_EQ, "!=": OP_NE, "<>": OP_NE, ">": OP_GT, ">=": OP_GE, "<": OP_LT, "<=": OP_LE}


class WhereOp(C.Structure):            # include/tostore_cuda.h: tsc_where_op
    _fields_ = [("kind", C.c_uint8), ("op", C.c_uint8), ("n", C.c_uint16),
                ("column_id", C.c_uint32), ("i_lo", C.c_int64), ("i_hi", C.c_int64),
                ("f_lo", C.c_double), ("f_hi", C.c_double), ("args_offset", C.c_uint32),
                ("reserved", C.c_uint32)]


def _dart_round(x: float) -> int:
    if x != x or math.isinf(x):
        raise ValueError("cannot convert NaN / infinity to an integer field operand")
    a = abs(x)
    r = a if a >= 2.0 ** 52 else (math.floor(a) + (1 if a - math.floor(a) >= 0.5 else 0))
    v = int(r) if x >= 0 else -int(r)
    return max(_INT64_MIN, min(_INT64_MAX, v))


# String.trim(): Unicode White_Space plus the byte-order mark (dart:core String.trim)
_DART_WS = frozenset([0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029,
                      0x202F, 0x205F, 0x3000, 0xFEFF] + list(range(0x2000, 0x200B)))


def dart_trim(s: str) -> str:
    a, b = 0, len(s)
    while a < b and ord(s[a]) in _DART_WS:
        a += 1
    while b > a and ord(s[b - 1]) in _DART_WS:
        b -= 1
    return s[a:b]


def convert_text(v) -> str:
    """`convertValue` for DataType.text (table_schema.dart:1421-1442): toString().trim()."""
    if isinstance(v, (bool, np.bool_)):
        v = "true" if v else "false"
    elif isinstance(v, (int, np.integer)):
        v = str(int(v))
    elif not isinstance(v, str):
        raise TypeError(f"operand {v!r} has no exact Dart toString() here (pass a string)")
    return dart_trim(v)


def convert_bool(v) -> int:
    """`convertValue` for DataType.boolean (table_schema.dart:1450-1459) as the 0 / 1 the
    device column holds; the boolean matcher orders false < true (value_matcher.dart:242-253)."""
    if isinstance(v, (bool, np.bool_)):
        return 1 if v else 0
    if isinstance(v, (int, np.integer)):
        return 1 if int(v) != 0 else 0
    if isinstance(v, (float, np.floating)):
        return 1 if float(v) != 0.0 else 0
    if isinstance(v, str):
        return 1 if v.lower() in ("true", "1", "yes") else 0
    raise TypeError(f"operand {v!r} cannot be converted for a boolean field")


def convert_datetime(v) -> str:
    """DataType.datetime is stored as `DateTime.toIso8601String()` and compared as a STRING
    (the datetime matcher is the text matcher, value_matcher.dart:211-240). A `datetime` is
    formatted the way Dart does (`yyyy-MM-ddTHH:mm:ss.mmm[uuu][Z]`: milliseconds always,
    microseconds only when non-zero, `Z` for UTC); a string is taken to be in that stored form
    already — the reference would re-parse it (`DateTime.parse`, table_schema.dart:1462-1475),
    which has no exact restatement here."""
    import datetime as _dt
    if isinstance(v, str):
        return v
    if isinstance(v, _dt.datetime):
        utc = False
        if v.tzinfo is not None:
            if v.utcoffset() != _dt.timedelta(0):
                raise ValueError("a Dart DateTime is local or UTC: convert the operand to UTC first")
            utc = True
        if not 0 <= v.year <= 9999:
            raise ValueError("year outside 0..9999")
        s = "%04d-%02d-%02dT%02d:%02d:%02d.%03d" % (v.year, v.month, v.day, v.hour, v.minute, v.second,
                                                    v.microsecond // 1000)
        if v.microsecond % 1000:
            s += "%03d" % (v.microsecond % 1000)
        return s + ("Z" if utc else "")
    raise TypeError(f"operand {v!r} cannot be converted for a datetime field")


def utf16_units(s: str) -> np.ndarray:
    """A string's UTF-16 code units (`String.codeUnits`); lone surrogates pass through."""
    return np.frombuffer(s.encode("utf-16-le", "surrogatepass"), dtype=np.uint16)


def utf16_pool(strings):
    """strings -> (units uint16 [total], offsets uint64 [n + 1]) as the C ABI takes them."""
    parts = [utf16_units(s) for s in strings]
    offsets = np.zeros(len(parts) + 1, dtype=np.uint64)
    if parts:
        offsets[1:] = np.cumsum([p.size for p in parts], dtype=np.uint64)
    units = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint16)
    if units.size == 0:
        units = np.zeros(1, dtype=np.uint16)      # a valid address for an empty pool
    return np.ascontiguousarray(units), offsets


def _convert(v, col_type: int):
    """`FieldSchema.convertValue` for integer / double / text fields
    (table_schema.dart:1371-1442)."""
    if col_type == COL_TEXT:
        return convert_text(v)
    if col_type == COL_BOOL:
        return convert_bool(v)
    if col_type == COL_DATETIME:
        return convert_datetime(v)
    if isinstance(v, (bool, np.bool_)):
        v = 1 if v else 0
    if isinstance(v, np.integer):
        v = int(v)
    if isinstance(v, np.floating):
        v = float(v)
    if col_type == COL_I64:
        if isinstance(v, int):
            if not _INT64_MIN <= v <= _INT64_MAX:
                raise ValueError(f"operand {v} does not fit an int64 field")
            return v
        if isinstance(v, float):
            return _dart_round(v)
    else:
        if isinstance(v, float):
            return v
        if isinstance(v, int):
            return float(v)
    raise TypeError(f"operand {v!r} is not numeric")


class WhereProgram:
    """Postfix program + IN-list values, ready for the C ABI."""

    def __init__(self):
        self.ops: List[WhereOp] = []
        self.args: List[Tuple[int, object]] = []     # (col_type, value); text: index into texts
        self.texts: List[str] = []                   # operand pool of the text leaves

    def _text(self, s: str) -> int:
        self.texts.append(s)
        return len(self.texts) - 1

    def _leaf(self, op: int, col: int = 0, col_type: int = COL_I64, lo=0, hi=0, n: int = 0,
              args_offset: int = 0) -> None:
        o = WhereOp(kind=W_LEAF, op=op, n=n, column_id=col, args_offset=args_offset)
        if col_type == COL_TEXT:                     # operands are named by pool index
            o.i_lo = self._text(lo) if isinstance(lo, str) else 0
            o.i_hi = self._text(hi) if isinstance(hi, str) else 0
        elif col_type == COL_I64:
            o.i_lo, o.i_hi = int(lo), int(hi)
        else:
            o.f_lo, o.f_hi = float(lo), float(hi)
        self.ops.append(o)

    def _node(self, kind: int, n: int) -> None:
        self.ops.append(WhereOp(kind=kind, n=n))

    def buffers(self):
        if len(self.ops) > MAX_OPS:
            raise ValueError(f"condition compiles to {len(self.ops)} steps (limit {MAX_OPS})")
        if len(self.args) > MAX_IN_ARGS:
            raise ValueError(f"IN lists hold {len(self.args)} values (limit {MAX_IN_ARGS})")
        ops = (WhereOp * max(len(self.ops), 1))(*self.ops)
        if len(self.texts) > MAX_TEXTS:
            raise ValueError(f"text operands: {len(self.texts)} (limit {MAX_TEXTS})")
        raw = np.zeros(max(len(self.args), 1), dtype=np.uint64)
        for i, (t, v) in enumerate(self.args):
            raw[i] = (np.array([v], dtype=np.int64) if t in (COL_I64, COL_TEXT)
                      else np.array([v], dtype=np.float64)).view(np.uint64)[0]
        return ops, len(self.ops), raw, len(self.args)

    def text_buffers(self):
        """(units, offsets) of the text operand pool for tsc_index_filter_where_text."""
        return utf16_pool(self.texts)


def compile_condition(cond: Dict[str, object], columns: Dict[str, Tuple[int, int]]) -> WhereProgram:
    """cond: map form of a `QueryCondition`; columns: field name -> (column_id, COL_*).
    An empty condition compiles to the empty program (matches every row)."""
    prog = WhereProgram()
    if cond:
        _emit(cond, columns, prog)
    return prog


def _emit(cond, columns, prog: WhereProgram) -> None:
    if not isinstance(cond, dict):
        raise TypeError("condition must be a map")
    if "AND" in cond or "OR" in cond:
        key = "AND" if "AND" in cond else "OR"
        kids = list(cond[key])
        if len(kids) > 63:
            raise ValueError("more than 63 children under one AND / OR")
        for c in kids:
            _emit(c, columns, prog)
        prog._node(W_AND if key == "AND" else W_OR, len(kids))
        return
    n_fields = 0
    for field, c in cond.items():                      # fields of one leaf: AND
        name = field.split(".")[-1] if field not in columns else field
        if name not in columns:
            raise KeyError(f"WHERE names field {field!r} which has no attribute column")
        col, t = columns[name]
        if t == COL_BOOL:
            c = _map_operands(c, convert_bool, "boolean")
            t = COL_I64
        elif t == COL_DATETIME:
            c = _map_operands(c, convert_datetime, "datetime")
            t = COL_TEXT
        if isinstance(c, dict):
            n_ops = 0
            for op, ov in c.items():                   # operators of one field: OR
                _emit_operator(op, ov, col, t, prog)
                n_ops += 1
            if n_ops != 1:
                prog._node(W_OR, n_ops) if n_ops else prog._leaf(OP_FALSE)
        elif c is None:
            prog._leaf(OP_IS_NULL, col, t)             # `condition == null -> value == null`
        else:
            prog._leaf(OP_EQ, col, t, _convert(c, t))
        n_fields += 1
    if n_fields != 1:
        prog._node(W_AND, n_fields)


def _map_operands(c, conv, what):
    """Operands of a condition on a host-mapped field (boolean -> 0 / 1, datetime -> ISO-8601
    string) through `conv`; None stays None."""
    cv = lambda x: None if x is None else conv(x)   # noqa: E731
    if not isinstance(c, dict):
        return cv(c)
    out = {}
    for op, ov in c.items():
        up = op.upper()
        if up == "BETWEEN" and isinstance(ov, dict) and "start" in ov and "end" in ov:
            out[op] = {"start": cv(ov["start"]), "end": cv(ov["end"])}
        elif up in ("IN", "NOT IN") and isinstance(ov, (list, tuple)):
            out[op] = [cv(x) for x in ov]
        elif up in ("LIKE", "NOT LIKE") and what == "boolean":
            raise NotImplementedError(f"{up} on a boolean field has no columnar GPU form")
        else:
            out[op] = cv(ov)
    return out


def _emit_operator(op: str, ov, col: int, t: int, prog: WhereProgram) -> None:
    up = op.upper()
    if up in _SIMPLE:
        if ov is None:
            # matcher(value, null): 0 iff value is null, else +1 (value_matcher.dart:160-163)
            code = {OP_EQ: OP_IS_NULL, OP_NE: OP_IS_NOT_NULL, OP_GT: OP_IS_NOT_NULL,
                    OP_GE: OP_IS_NOT_NULL, OP_LT: OP_FALSE, OP_LE: OP_FALSE}[_SIMPLE[up]]
            prog._leaf(code, col, t)
        else:
            prog._leaf(_SIMPLE[up], col, t, _convert(ov, t))
    elif up == "BETWEEN":
        if not isinstance(ov, dict) or "start" not in ov or "end" not in ov:
            prog._leaf(OP_FALSE)
        else:
            prog._leaf(OP_BETWEEN, col, t, _convert(ov["start"], t), _convert(ov["end"], t))
    elif up in ("IN", "NOT IN"):
        if not isinstance(ov, (list, tuple)):
            prog._leaf(OP_FALSE if up == "IN" else OP_TRUE)
        else:
            vals = [_convert(x, t) for x in ov if x is not None]
            off = len(prog.args)
            if t == COL_TEXT:
                vals = [prog._text(v) for v in vals]
            prog.args.extend((t, v) for v in vals)
            prog._leaf(OP_IN if up == "IN" else OP_NOT_IN, col, t, n=len(vals), args_offset=off)
    elif up == "IS":
        prog._leaf(OP_IS_NULL if ov is None else OP_FALSE, col, t)
    elif up == "IS NOT":
        prog._leaf(OP_IS_NOT_NULL if ov is None else OP_FALSE, col, t)
    elif up in ("LIKE", "NOT LIKE"):
        if t != COL_TEXT:
            raise NotImplementedError(f"{up} on a numeric field has no columnar GPU form")
        if ov is None:
            prog._leaf(OP_FALSE)                        # `compareValue is! String` (:599-604)
        else:
            prog._leaf(OP_LIKE if up == "LIKE" else OP_NOT_LIKE, col, t, _convert(ov, t))
    else:
        raise NotImplementedError(f"operator {op!r} has no columnar GPU form")


class QueryCondition:
    """Small builder for the map form (`QueryCondition.where / or / build`,
    query/query_condition.dart:117-260): `where` ANDs onto the current group, `orWhere`
    starts an alternative."""

    def __init__(self):
        self._groups: List[List[dict]] = [[]]

    def where(self, field: str, operator, value=None) -> "QueryCondition":
        self._groups[-1].append(_build(field, operator, value))
        return self

    def orWhere(self, field: str, operator, value=None) -> "QueryCondition":
        self._groups.append([_build(field, operator, value)])
        return self

    def whereIn(self, field, values):
        return self.where(field, "IN", list(values))

    def whereNotIn(self, field, values):
        return self.where(field, "NOT IN", list(values))

    def whereBetween(self, field, start, end):
        return self.where(field, "BETWEEN", [start, end])

    def whereNull(self, field):
        return self.where(field, "IS", None)

    def whereNotNull(self, field):
        return self.where(field, "IS NOT", None)

    # convenience forms, one `where` each (query_condition.dart:574-678)
    def whereLike(self, field, pattern):
        return self.where(field, "LIKE", pattern)

    def whereNotLike(self, field, pattern):
        return self.where(field, "NOT LIKE", pattern)

    def whereContains(self, field, value):
        return self.where(field, "LIKE", f"%{value}%")

    def whereNotContains(self, field, value):
        return self.where(field, "NOT LIKE", f"%{value}%")

    def whereStartsWith(self, field, prefix):
        return self.where(field, "LIKE", f"{prefix}%")

    def whereEndsWith(self, field, suffix):
        return self.where(field, "LIKE", f"%{suffix}")

    def whereEqual(self, field, value):
        return self.where(field, "=", value)

    def whereNotEqual(self, field, value):
        return self.where(field, "!=", value)

    def whereGreaterThan(self, field, value):
        return self.where(field, ">", value)

    def whereGreaterThanOrEqualTo(self, field, value):
        return self.where(field, ">=", value)

    def whereLessThan(self, field, value):
        return self.where(field, "<", value)

    def whereLessThanOrEqualTo(self, field, value):
        return self.where(field, "<=", value)

    def whereTrue(self, field):
        return self.where(field, "=", True)

    def whereFalse(self, field):
        return self.where(field, "=", False)

    def whereNotEmpty(self, field):
        return self.whereNotNull(field).where(field, "!=", "")

    def condition(self, other: "QueryCondition") -> "QueryCondition":
        """AND a whole sub-condition onto the current group (`condition`, :290-367)."""
        if not other.isEmpty:
            self._groups[-1].append(other.build())
        return self

    def whereEmpty(self, field):
        """NULL or the empty string (:659-663)."""
        return self.condition(QueryCondition().whereNull(field).orWhere(field, "=", ""))

    def whereContainsAny(self, field, values):
        """LIKE '%v%' for any of the values (:585-597)."""
        sub = QueryCondition()
        for i, v in enumerate(values):
            (sub.orWhere if i else sub.where)(field, "LIKE", f"%{v}%")
        return self.condition(sub)

    @property
    def isEmpty(self) -> bool:
        return not any(self._groups)

    def build(self) -> dict:
        groups = [g for g in self._groups if g]
        if not groups:
            return {}
        ands = [g[0] if len(g) == 1 else {"AND": g} for g in groups]
        return ands[0] if len(ands) == 1 else {"OR": ands}


def _build(field, operator, value):
    """`_buildCondition` (query_condition.dart:478-520)."""
    if value is None:
        if operator in ("IS", "IS NOT"):
            return {field: {operator: None}}
        return {field: operator}
    op = str(operator).upper()
    if op == "BETWEEN":
        return {field: {"BETWEEN": {"start": value[0], "end": value[1]}}}
    return {field: {op: value}}
