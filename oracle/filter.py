"""CPU oracle of the reference's metadata filter evaluation -- TEST INFRASTRUCTURE ONLY (see oracle/syzgy_oracle.c).

Restates, for syntax trees given as nested tuples, what query/compiler.go does to one document:
CreateFilterFunction (477-497), CompileExpression (15-165), evaluateOperation (167-264), compareValues (266-326),
evaluateFunction (328-363), evaluateIn / Contains / StartsWith / EndsWith / Matches (379-426), getField (428-444), and
BuildFilter's "an error means false" (collection.go:204-218).

Node spelling (mirrors query/parser.go's node types):
    ("ident", name)                       IdentifierNode
    ("value", v)                          ValueNode: float, str, bool or None (parser.go:185-199, 472-489)
    ("expr", op, left, right)             ExpressionNode; NOT has left = None
    ("array", [nodes])                    ArrayNode
    ("func", name, [nodes])               FunctionNode: EXISTS, DOES_NOT_EXIST, LENGTH

Parity pinning: tests/test_filter.py replays the syntax-tree cases of the reference's own TestCompileExpression
(query/compiler_test.go:10-186) and the flat cases of TestCreateFilterFunction (196-318, lowered by hand since the
lexer/parser stay in Go) against this file before using it to judge the device path.  MATCHES uses Python's `re`
instead of Go's RE2 (the syntaxes agree on the patterns the reference tests use).
"""
import json
import re


class EvalError(Exception):
    """An `error` return of the Go code."""


def parse_metadata(raw: bytes):
    """json.Unmarshal into interface{}: every number is a float64; returns (ok, value)."""
    def no_const(name):
        raise ValueError(name)
    try:
        return True, json.loads(raw.decode("utf-8"), parse_int=float, parse_constant=no_const)
    except Exception:
        return False, None


def _deep_equal(a, b) -> bool:
    # reflect.DeepEqual over interface{} values from JSON and from literals
    if isinstance(a, bool) or isinstance(b, bool):
        return isinstance(a, bool) and isinstance(b, bool) and a == b
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, float) and isinstance(b, float):
        return a == b
    if isinstance(a, str) and isinstance(b, str):
        return a == b
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(_deep_equal(x, y) for x, y in zip(a, b))
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_deep_equal(a[k], b[k]) for k in a)
    return False


def _compare(op, left, right) -> bool:  # compiler.go:266-326
    if isinstance(left, int) and not isinstance(left, bool):  # reflect.Int (only LENGTH produces one): toInt64(right)
        if isinstance(right, bool) or not isinstance(right, (int, float)):
            raise EvalError("cannot convert to int64")
        l, r = left, int(right)
    elif isinstance(left, float) and not isinstance(left, bool):
        if not isinstance(right, float) or isinstance(right, bool):
            raise EvalError("cannot convert to float64")
        l, r = left, right
    elif isinstance(left, str):
        if not isinstance(right, str):
            raise EvalError("cannot compare string with non-string")
        l, r = left.encode("utf-8"), right.encode("utf-8")  # Go compares strings bytewise
    else:
        raise EvalError("unsupported comparison")
    return {">": l > r, ">=": l >= r, "<": l < r, "<=": l <= r}[op]


def _get_field(data, path):  # compiler.go:428-444 (returns at the first key)
    current = data
    for key in path:
        if isinstance(current, dict):
            return current.get(key)
        if isinstance(current, list):
            if key == "*":
                return current
            raise EvalError("cannot use dot notation on array")
        raise EvalError("cannot access field")
    return current


def _strings(left, right, what):
    if not isinstance(left, str) or not isinstance(right, str):
        raise EvalError(f"{what} operation requires string operands")
    return left, right


def _operation(op, left, right):  # compiler.go:167-264
    if op == "==":
        return _deep_equal(left, right)
    if op == "!=":
        return not _deep_equal(left, right)
    if op in (">", ">=", "<", "<="):
        return _compare(op, left, right)
    if op == "AND":
        if not isinstance(left, bool) or not isinstance(right, bool):
            raise EvalError("AND operation requires boolean operands")
        return left and right
    if op == "OR":
        if not isinstance(left, bool):
            raise EvalError("OR operation requires boolean operands")
        if left:
            return True
        if not isinstance(right, bool):
            raise EvalError("OR operation requires boolean operands")
        return right
    if op == "NOT":
        if not isinstance(right, bool):
            raise EvalError("NOT operation requires a boolean operand")
        return not right
    if op in ("IN", "NOT_IN"):
        if not isinstance(right, list):
            raise EvalError("IN operator requires a list on the right side")
        found = any(_deep_equal(left, item) for item in right)
        return found if op == "IN" else not found
    if op == "CONTAINS":
        l, r = _strings(left, right, op)
        return r in l
    if op == "STARTS_WITH":
        l, r = _strings(left, right, op)
        return l.startswith(r)
    if op == "ENDS_WITH":
        l, r = _strings(left, right, op)
        return l.endswith(r)
    if op == "MATCHES":
        l, r = _strings(left, right, op)
        return re.search(r, l) is not None
    if op == ".":
        if isinstance(left, dict):
            if right not in left:
                raise EvalError("key not found in map")
            return left[right]
        if isinstance(left, list):
            if right == "length":
                return float(len(left))
            raise EvalError("invalid operation on array")
        raise EvalError("left operand of '.' must be a map or array")
    raise EvalError(f"unsupported operator {op}")


def evaluate(node, data):  # compiler.go:15-165
    kind = node[0]
    if kind == "value":
        return node[1]
    if kind == "ident":
        return _get_field(data, node[1].split("."))
    if kind == "array":
        return [evaluate(e, data) for e in node[1]]
    if kind == "expr":
        _, op, left, right = node
        lval = evaluate(left, data) if left is not None else None  # `case nil` compiles to nil
        if op == ".":
            if right[0] != "ident":
                raise EvalError("right side of '.' must be an identifier")
            rval = right[1]
        else:
            rval = evaluate(right, data)
        return _operation(op, lval, rval)
    if kind == "func":
        _, name, args = node
        if name == "DOES_NOT_EXIST":
            if len(args) != 1 or args[0][0] != "ident":
                raise EvalError("DOES_NOT_EXIST function argument must be an identifier")
            return (args[0][1] not in data) if isinstance(data, dict) else False
        if name == "EXISTS":
            if len(args) != 1:
                raise EvalError("EXISTS function requires exactly one argument")
            try:
                evaluate(args[0], data)
                return True
            except EvalError:
                return False
        if name == "LENGTH":
            arg = evaluate(args[0], data)
            if isinstance(arg, (str, list, dict)):
                return len(arg.encode("utf-8")) if isinstance(arg, str) else len(arg)  # a Go int, not a float64
            raise EvalError("LENGTH function not supported")
        raise EvalError(f"unsupported function {name}")
    raise EvalError(f"unsupported node {kind}")


def filter_document(node, raw: bytes) -> bool:
    """BuildFilter(query)(id, metadata): errors and non-boolean results are false."""
    ok, data = parse_metadata(raw)
    if not ok:
        return False
    try:
        result = evaluate(node, data)
    except EvalError:
        return False
    return result is True
