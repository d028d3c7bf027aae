"""Host side of the device filter path: what the Go shim does around szg_meta_upsert / szg_filter_mask
(INTEGRATION.md shows the Go spelling).  Two jobs:

* `column_values`: one pass over a document's metadata JSON at AddDocument / UpdateDocument time, producing the
  document kind and one typed value per mirrored field (top-level keys, or dotted paths evaluated with the '.'
  operator's rules, query/compiler.go:222-244);
* `lower`: a filter's syntax tree (the node types of query/parser.go, spelled as tuples -- see `oracle/filter.py` for
  the same spelling on the checker's side) to the postfix program szg_filter_mask runs.  Trees outside the covered
  subset raise `Unsupported`; the caller then evaluates the Go predicate per document and uses szg_mask_create, as
  before.

Nothing here evaluates a filter on the CPU: the program runs on the device only.
"""
import json
import re

from . import _capi

MV_MISSING, MV_NULL, MV_BOOL, MV_NUMBER, MV_STRING, MV_OTHER, MV_ERROR = range(7)
DOC_INVALID, DOC_OBJECT, DOC_OTHER = range(3)

_CMP = {"==": _capi.FOP_EQ, "!=": _capi.FOP_NE, "<": _capi.FOP_LT, "<=": _capi.FOP_LE, ">": _capi.FOP_GT, ">=": _capi.FOP_GE}
_STR = {"CONTAINS": _capi.FOP_CONTAINS, "STARTS_WITH": _capi.FOP_STARTS_WITH, "ENDS_WITH": _capi.FOP_ENDS_WITH}


class Unsupported(Exception):
    """The tree uses something the device program does not cover."""


def _typed(v):
    if v is None:
        return (MV_NULL, None)
    if isinstance(v, bool):
        return (MV_BOOL, v)
    if isinstance(v, (int, float)):
        return (MV_NUMBER, float(v))
    if isinstance(v, str):
        return (MV_STRING, v)
    return (MV_OTHER, None)


def _field(doc: dict, path):
    # first hop = getField on the document (a missing key reads as nil); further hops = the '.' operator
    if path[0] not in doc:
        if len(path) == 1:
            return (MV_MISSING, None)
        return (MV_ERROR, None)  # '.' on nil
    cur = doc[path[0]]
    for key in path[1:]:
        if isinstance(cur, dict):
            if key not in cur:
                return (MV_ERROR, None)
            cur = cur[key]
        elif isinstance(cur, list):
            if key != "length":
                return (MV_ERROR, None)
            cur = float(len(cur))
        else:
            return (MV_ERROR, None)
    return _typed(cur)


def column_values(raw: bytes, fields):
    """(document kind, [(kind, value) per field]) of one metadata blob; `fields` are key names or dotted paths."""
    def no_const(name):
        raise ValueError(name)
    try:
        doc = json.loads(raw.decode("utf-8"), parse_int=float, parse_constant=no_const)
    except Exception:
        return DOC_INVALID, [(MV_MISSING, None)] * len(fields)
    if not isinstance(doc, dict):
        return DOC_OTHER, [(MV_MISSING, None)] * len(fields)
    return DOC_OBJECT, [_field(doc, f.split(".")) for f in fields]


def _path(node):
    """IdentifierNode, or a chain of '.' over identifiers, as a dotted field name (else None)."""
    if node[0] == "ident":
        return node[1]
    if node[0] == "expr" and node[1] == "." and node[3][0] == "ident":
        left = _path(node[2])
        return None if left is None else left + "." + node[3][1]
    return None


def lower(node, columns, dictionary=None):
    """Postfix program (list of op dicts for Index.filter_mask) of a syntax tree.  `columns`: field name -> column id.
    `dictionary`: callable returning the collection's string dictionary (needed by MATCHES only)."""
    out = []

    def operand(n):
        p = _path(n)
        if p is not None:
            if p not in columns:
                raise Unsupported(f"field {p} is not mirrored")
            out.append({"op": _capi.FOP_COL, "arg": columns[p]})
            return "col"
        if n[0] == "value":
            v = n[1]
            if v is None:
                out.append({"op": _capi.FOP_NULL})
            elif isinstance(v, bool):
                out.append({"op": _capi.FOP_BOOL, "num": 1.0 if v else 0.0})
            elif isinstance(v, (int, float)):
                out.append({"op": _capi.FOP_NUM, "num": float(v)})
            elif isinstance(v, str):
                out.append({"op": _capi.FOP_STR, "str": v})
            else:
                raise Unsupported("literal type")
            return "lit"
        if n[0] not in ("expr", "func"):
            raise Unsupported(f"node {n[0]}")
        expr(n)
        return "expr"

    def expr(n):
        kind = n[0]
        if kind == "func":
            _, name, args = n
            if name in ("EXISTS", "DOES_NOT_EXIST") and len(args) == 1 and args[0][0] == "ident" and "." not in args[0][1]:
                if args[0][1] not in columns:
                    raise Unsupported(f"field {args[0][1]} is not mirrored")
                out.append({"op": _capi.FOP_EXISTS if name == "EXISTS" else _capi.FOP_NOT_EXISTS, "arg": columns[args[0][1]]})
                return
            raise Unsupported(f"function {name}")
        if kind != "expr":
            if kind not in ("ident", "value"):
                raise Unsupported(f"node {kind}")
            operand(n)  # a bare field or literal: passes only where it is a bool
            return
        _, op, left, right = n
        if op in _CMP:
            a, b = operand(left), operand(right)
            if a == "col" and b == "col":
                raise Unsupported("comparison between two fields")
            out.append({"op": _CMP[op]})
        elif op in ("AND", "OR"):
            operand(left)
            operand(right)
            out.append({"op": _capi.FOP_AND if op == "AND" else _capi.FOP_OR})
        elif op == "NOT":
            operand(right)
            out.append({"op": _capi.FOP_NOT})
        elif op in ("IN", "NOT_IN"):
            if right[0] != "array" or any(e[0] != "value" for e in right[1]):
                raise Unsupported("IN needs a list of literals")
            operand(left)
            for e in right[1]:
                operand(e)
            out.append({"op": _capi.FOP_IN if op == "IN" else _capi.FOP_NOT_IN, "arg": len(right[1])})
        elif op in _STR or op == "MATCHES":
            if right[0] != "value" or not isinstance(right[1], str):
                raise Unsupported(f"{op} needs a string literal on the right")
            if operand(left) == "lit":
                raise Unsupported(f"{op} of a literal")
            if op == "MATCHES":
                if dictionary is None:
                    raise Unsupported("MATCHES needs the dictionary")
                rx = re.compile(right[1])  # the Go shim uses regexp.MatchString here (compiler.go:422-426)
                out.append({"op": _capi.FOP_STR_TABLE, "table": bytes(1 if rx.search(s) else 0 for s in dictionary())})
            else:
                out.append({"op": _STR[op], "str": right[1]})
        elif op == ".":
            if _path(n) is None:
                raise Unsupported("'.' on something that is not a field path")
            operand(n)
        else:
            raise Unsupported(f"operator {op}")

    expr(node)
    return out
