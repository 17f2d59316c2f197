"""CPU restatement of the reference's structured-condition evaluation for integer,
double and text fields — TEST INFRASTRUCTURE ONLY (checker for tsc_index_filter_where; never imported
by the product path).

Follows, relative to /root/reference/lib/src:
  * operand normalisation to the field's type
      query/query_condition.dart:743-815 (normalize / _convertConditionValue)
      model/table_schema.dart:1356-1421  (convertValue: integer <- double.round(),
                                          double <- int.toDouble())
  * tree walk        handler/value_matcher.dart:476-511 (_matchNode / _matchAllConditions)
  * per-field match  handler/value_matcher.dart:513-568 (_matchFieldCondition /
                     _matchSingleCondition: an operator map is an OR of its operators)
  * operators        handler/value_matcher.dart:570-612 (_evaluateOperator)
  * numeric order    handler/value_matcher.dart:150-174 -> Dart num.compareTo
                     (-0.0 < 0.0; NaN above everything and equal to itself)
  * text fields      model/table_schema.dart:1421-1442 (convertValue: toString().trim()),
                     handler/value_matcher.dart:211-240 (String.compareTo = UTF-16 code-unit
                     order), :318-331 (matchesLike: LIKE pattern -> anchored RegExp, `%` ->
                     `.*`, `_` -> `.`, everything else escaped; `.` of a RegExp without dotAll
                     / unicode is one code unit other than \n \r U+2028 U+2029), :599-604
                     (LIKE and NOT LIKE are both false on NULL / a non-string pattern)
Parity is unpinned by the reference (it has no test for any of this and no WHERE for
vectors); pinned here by hand-written known answers in tests/test_where.py.

Pure Python loops over records: for small cases. `evaluate_columns` is the same walk
driven from column arrays.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional, Sequence

INT64_MIN, INT64_MAX = -(1 << 63), (1 << 63) - 1


# ---- Dart arithmetic helpers ---------------------------------------------------------
def dart_compare(a, b) -> int:
    """num.compareTo for int/int and double/double (and, exactly, for mixed pairs)."""
    a_nan = isinstance(a, float) and a != a
    b_nan = isinstance(b, float) and b != b
    if a_nan or b_nan:
        return 0 if (a_nan and b_nan) else (1 if a_nan else -1)
    if a < b:
        return -1
    if a > b:
        return 1
    if a == 0:
        a_neg = isinstance(a, float) and math.copysign(1.0, a) < 0
        b_neg = isinstance(b, float) and math.copysign(1.0, b) < 0
        if a_neg != b_neg:
            return -1 if a_neg else 1
    return 0


def dart_round(x: float) -> int:
    """double.round(): half away from zero; clamps to int64 like the VM's toInt()."""
    if x != x or x in (math.inf, -math.inf):
        raise ValueError("double.round() of NaN / infinity throws UnsupportedError in Dart")
    a = abs(x)
    r = a if a >= 2.0 ** 52 else (math.floor(a) + (1 if a - math.floor(a) >= 0.5 else 0))
    v = int(r) if x >= 0 else -int(r)
    return max(INT64_MIN, min(INT64_MAX, v))


# ---- Dart strings ----------------------------------------------------------------------
_TRIM = set([0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F,
             0x205F, 0x3000, 0xFEFF]) | set(range(0x2000, 0x200B))
_NOT_LT = "[^\n\r\u2028\u2029]"


def code_units(s: str) -> List[int]:
    """String.codeUnits (UTF-16; lone surrogates pass through)."""
    b = s.encode("utf-16-le", "surrogatepass")
    return [b[i] | (b[i + 1] << 8) for i in range(0, len(b), 2)]


def dart_string_compare(a: str, b: str) -> int:
    """String.compareTo: lexicographic over UTF-16 code units."""
    ua, ub = code_units(a), code_units(b)
    return -1 if ua < ub else (1 if ua > ub else 0)


def dart_trim(s: str) -> str:
    u = list(s)
    while u and ord(u[0]) in _TRIM:
        u.pop(0)
    while u and ord(u[-1]) in _TRIM:
        u.pop()
    return "".join(u)


def matches_like(value: str, pattern: str) -> bool:
    """ValueMatcher.matchesLike (value_matcher.dart:318-331) through Python's regex engine,
    one regex character per UTF-16 code unit."""
    rx = "".join(_NOT_LT + "*" if c == 0x25 else _NOT_LT if c == 0x5F else re.escape(chr(c))
                 for c in code_units(pattern))
    return re.fullmatch(rx, "".join(chr(c) for c in code_units(value)), flags=re.S) is not None


def convert_value(v, col_type: str):
    """FieldSchema.convertValue for DataType.integer / double / text operands."""
    if v is None:
        return None
    if col_type == "bool":        # table_schema.dart:1450-1459
        if isinstance(v, bool):
            return v
        if isinstance(v, int):
            return v != 0
        if isinstance(v, float):
            return v != 0.0
        if isinstance(v, str):
            return v.lower() in ("true", "1", "yes")
        raise TypeError(f"unsupported operand {v!r} for a boolean field")
    if col_type == "text":
        if isinstance(v, bool):
            v = "true" if v else "false"
        elif isinstance(v, int):
            v = str(v)
        if not isinstance(v, str):
            raise TypeError(f"unsupported operand {v!r} for a text field")
        return dart_trim(v)
    if isinstance(v, bool):
        v = 1 if v else 0
    if col_type == "i64":
        if isinstance(v, int):
            return v
        if isinstance(v, float):
            return dart_round(v)
        raise TypeError(f"unsupported operand {v!r} for an integer field")
    if isinstance(v, float):
        return v
    if isinstance(v, int):
        return float(v)          # int.toDouble(): round to nearest
    raise TypeError(f"unsupported operand {v!r} for a double field")


def _matcher(a, b) -> int:
    """Nullable numeric matcher (value_matcher.dart:160-174)."""
    if a is None or b is None:
        return 0 if a is b else (-1 if a is None else 1)
    if isinstance(a, str) and isinstance(b, str):      # text matcher (:225-240)
        return dart_string_compare(a, b)
    if isinstance(a, bool) and isinstance(b, bool):    # boolean matcher (:242-253): false < true
        return 0 if a == b else (1 if a else -1)
    return dart_compare(a, b)


def evaluate_operator(value, op: str, cmp):
    """_evaluateOperator (value_matcher.dart:570-612), numeric operators only."""
    op = op.upper()
    if op == "=":
        return _matcher(value, cmp) == 0
    if op in ("!=", "<>"):
        return _matcher(value, cmp) != 0
    if op == ">":
        return value is not None and _matcher(value, cmp) > 0
    if op == ">=":
        return value is not None and _matcher(value, cmp) >= 0
    if op == "<":
        return value is not None and _matcher(value, cmp) < 0
    if op == "<=":
        return value is not None and _matcher(value, cmp) <= 0
    if op == "IN":
        if value is None or not isinstance(cmp, (list, tuple)):
            return False
        return any(_matcher(value, x) == 0 for x in cmp)
    if op == "NOT IN":
        if value is None or not isinstance(cmp, (list, tuple)):
            return True
        return not any(_matcher(value, x) == 0 for x in cmp)
    if op == "BETWEEN":
        if value is None or not isinstance(cmp, dict) or "start" not in cmp or "end" not in cmp:
            return False
        return _matcher(value, cmp["start"]) >= 0 and _matcher(value, cmp["end"]) <= 0
    if op == "LIKE":
        if value is None or not isinstance(cmp, str):
            return False
        return matches_like(str(value), cmp)
    if op == "NOT LIKE":
        if value is None or not isinstance(cmp, str):
            return False
        return not matches_like(str(value), cmp)
    if op == "IS":
        return value is None and cmp is None
    if op == "IS NOT":
        return value is not None and cmp is None
    raise ValueError(f"unknown operator {op!r}")


def normalize_condition(cond, col_types: Dict[str, str]):
    """QueryCondition.normalize restricted to numeric fields (operands -> field type).
    One deliberate difference: a NULL operand stays NULL (the reference maps it through
    getDefaultValue(), table_schema.dart:1358-1360, which can turn `IS NULL` into a
    comparison with the field's default; documented in DESIGN.md)."""
    if not isinstance(cond, dict):
        raise TypeError("condition must be a map")
    out = {}
    for key, val in cond.items():
        if key in ("AND", "OR"):
            out[key] = [normalize_condition(c, col_types) for c in val]
            continue
        t = col_types[key]
        if isinstance(val, dict):
            m = {}
            for op, ov in val.items():
                up = op.upper()
                if up == "BETWEEN" and isinstance(ov, dict):
                    m[op] = {"start": convert_value(ov["start"], t), "end": convert_value(ov["end"], t)}
                elif up in ("IN", "NOT IN") and isinstance(ov, (list, tuple)):
                    m[op] = [convert_value(x, t) for x in ov]
                else:
                    m[op] = convert_value(ov, t)
            out[key] = m
        else:
            out[key] = convert_value(val, t)
    return out


def match_record(cond, record: Dict[str, object]) -> bool:
    """_matchNode / _matchAllConditions on the map form of a condition tree."""
    if "AND" in cond:
        return all(match_record(c, record) for c in cond["AND"])
    if "OR" in cond:
        kids = cond["OR"]
        return True if not kids else any(match_record(c, record) for c in kids)
    for field, c in cond.items():
        value = record.get(field)
        if isinstance(c, dict):
            if not any(evaluate_operator(value, op, ov) for op, ov in c.items()):
                return False
        elif c is None:
            if value is not None:
                return False
        elif _matcher(value, c) != 0:
            return False
    return True


def evaluate_columns(cond, columns: Dict[str, Sequence], col_types: Dict[str, str],
                     n_rows: Optional[int] = None) -> List[bool]:
    """Row-by-row evaluation over column arrays (None = NULL); rows beyond a column's end
    are NULL. Returns one bool per row."""
    if n_rows is None:
        n_rows = max((len(v) for v in columns.values()), default=0)
    norm = normalize_condition(cond, col_types) if cond else {}
    out = []
    for r in range(n_rows):
        rec = {}
        for name, vals in columns.items():
            v = vals[r] if r < len(vals) else None
            if v is not None:
                t = col_types[name]
                v = (int(v) if t == "i64" else str(v) if t == "text" else bool(v) if t == "bool"
                     else float(v))
            rec[name] = v
        out.append(match_record(norm, rec) if norm else True)
    return out
