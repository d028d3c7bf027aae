"""Metadata filters: the oracle (oracle/filter.py) against the reference's own golden cases, the host-side lowering,
and -- on the GPU -- the device program against the oracle, document by document."""
import json
import random

import numpy as np
import pytest

from oracle import filter as of
from syzgydb_b200 import _capi
from syzgydb_b200 import filter as hf

I, V, E, A, F = (lambda n: ("ident", n)), (lambda v: ("value", v)), (lambda op, l, r: ("expr", op, l, r)), \
    (lambda *e: ("array", list(e))), (lambda n, *a: ("func", n, list(a)))

# query/compiler_test.go:10-186 TestCompileExpression: (name, syntax tree, document, expected)
REFERENCE_TREE_CASES = [
    ("Simple Equality", E("==", I("age"), V(30.0)), '{"age": 30}', True),
    ("Simple Inequality", E("!=", I("age"), V(25.0)), '{"age": 30}', True),
    ("Greater Than", E(">", I("age"), V(25.0)), '{"age": 30}', True),
    ("Less Than or Equal", E("<=", I("age"), V(30.0)), '{"age": 30}', True),
    ("Logical AND", E("AND", E(">", I("age"), V(25.0)), E("==", I("status"), V("active"))), '{"age": 30, "status": "active"}', True),
    ("Logical OR", E("OR", E("<", I("age"), V(25.0)), E("==", I("status"), V("active"))), '{"age": 30, "status": "active"}', True),
    ("Logical NOT", E("NOT", None, E("==", I("status"), V("inactive"))), '{"status": "active"}', True),
    ("IN Operator", E("IN", I("status"), A(V("active"), V("pending"))), '{"status": "active"}', True),
    ("NOT IN Operator", E("NOT_IN", I("status"), A(V("inactive"), V("pending"))), '{"status": "active"}', True),
    ("CONTAINS Operator", E("CONTAINS", I("description"), V("urgent")), '{"description": "This is an urgent message"}', True),
    ("STARTS_WITH Operator", E("STARTS_WITH", I("filename"), V("report_")), '{"filename": "report_2023.pdf"}', True),
    ("ENDS_WITH Operator", E("ENDS_WITH", I("email"), V("@example.com")), '{"email": "user@example.com"}', True),
    ("MATCHES Operator", E("MATCHES", I("username"), V("^[a-z0-9_]{3,16}$")), '{"username": "john_doe123"}', True),
    ("EXISTS Function", F("EXISTS", I("optional_field")), '{"optional_field": "value"}', True),
    ("DOES_NOT_EXIST Function", F("DOES_NOT_EXIST", I("optional_field")), '{"other_field": "value"}', True),
    ("LENGTH Function", E(">=", F("LENGTH", I("tags")), V(3.0)), '{"tags": ["red", "green", "blue", "yellow"]}', True),
]
DOT = lambda a, b: E(".", I(a), I(b))
# query/compiler_test.go:196-318 TestCreateFilterFunction, the cases without array indexing, lowered by hand to the
# trees query/parser.go builds for them (the lexer and parser stay in Go)
REFERENCE_QUERY_CASES = [
    ("age == 30", E("==", I("age"), V(30.0)), '{"age": 30}', True),
    ("(age >= 18 AND status == 'active') OR role == 'admin'",
     E("OR", E("AND", E(">=", I("age"), V(18.0)), E("==", I("status"), V("active"))), E("==", I("role"), V("admin"))),
     '{"age": 25, "status": "active", "role": "user"}', True),
    ("user.email ENDS_WITH '@example.com'", E("ENDS_WITH", DOT("user", "email"), V("@example.com")),
     '{"user": {"email": "john@example.com"}}', True),
    ("status IN ['active', 'pending']", E("IN", I("status"), A(V("active"), V("pending"))), '{"status": "pending"}', True),
    ("status NOT IN ['inactive', 'suspended']", E("NOT_IN", I("status"), A(V("inactive"), V("suspended"))), '{"status": "active"}', True),
    ("(user.age > 25 AND (user.status == 'active' OR user.role == 'admin')) AND company.name STARTS_WITH 'Tech'",
     E("AND", E("AND", E(">", DOT("user", "age"), V(25.0)),
                E("OR", E("==", DOT("user", "status"), V("active")), E("==", DOT("user", "role"), V("admin")))),
       E("STARTS_WITH", DOT("company", "name"), V("Tech"))),
     '{"user": {"age": 30, "status": "inactive", "role": "admin"}, "company": {"name": "TechCorp"}}', True),
    ("name CONTAINS 'John' AND email ENDS_WITH '@example.com' AND id STARTS_WITH 'USER'",
     E("AND", E("AND", E("CONTAINS", I("name"), V("John")), E("ENDS_WITH", I("email"), V("@example.com"))),
       E("STARTS_WITH", I("id"), V("USER"))),
     '{"name": "John Doe", "email": "johndoe@example.com", "id": "USER123"}', True),
    ("price > 100 AND price < 200 AND quantity >= 5 AND discount <= 0.2",
     E("AND", E("AND", E("AND", E(">", I("price"), V(100.0)), E("<", I("price"), V(200.0))), E(">=", I("quantity"), V(5.0))),
       E("<=", I("discount"), V(0.2))), '{"price": 150, "quantity": 10, "discount": 0.15}', True),
    ("is_active == true AND is_deleted == false", E("AND", E("==", I("is_active"), V(True)), E("==", I("is_deleted"), V(False))),
     '{"is_active": true, "is_deleted": false}', True),
    ("optional_field == NULL AND required_field != NULL",
     E("AND", E("==", I("optional_field"), V(None)), E("!=", I("required_field"), V(None))), '{"required_field": "value"}', True),
    ("username MATCHES '^[a-z0-9_]{3,16}$'", E("MATCHES", I("username"), V("^[a-z0-9_]{3,16}$")), '{"username": "john_doe123"}', True),
    ("((a > 10 OR b < 5) AND (c == true OR d != false)) OR (e IN [1, 2, 3] AND f NOT IN ['x', 'y', 'z'])",
     E("OR", E("AND", E("OR", E(">", I("a"), V(10.0)), E("<", I("b"), V(5.0))),
               E("OR", E("==", I("c"), V(True)), E("!=", I("d"), V(False)))),
       E("AND", E("IN", I("e"), A(V(1.0), V(2.0), V(3.0))), E("NOT_IN", I("f"), A(V("x"), V("y"), V("z"))))),
     '{"a": 15, "b": 7, "c": false, "d": true, "e": 2, "f": "w"}', True),
]


@pytest.mark.parametrize("name,tree,doc,want", REFERENCE_TREE_CASES + REFERENCE_QUERY_CASES, ids=lambda x: x if isinstance(x, str) and len(x) < 40 else None)
def test_oracle_reproduces_the_references_golden_cases(name, tree, doc, want):
    assert of.filter_document(tree, doc.encode()) is want


def test_oracle_error_and_type_rules():
    d = b'{"age": 30, "name": "bob", "tags": [1, 2], "n": null, "ok": true}'
    f = of.filter_document
    assert not f(E(">", I("missing"), V(1.0)), d)                       # nil on the left of a comparison: error
    assert not f(E("NOT", None, E(">", I("missing"), V(1.0))), d)       # ... which poisons the whole expression
    assert not f(E("OR", E("==", I("age"), V(30.0)), E(">", I("name"), V(1.0))), d)  # operands are evaluated before OR
    assert f(E("OR", E("==", I("age"), V(30.0)), I("name")), d)         # a true left operand hides a non-bool right one
    assert not f(E("OR", E("==", I("age"), V(31.0)), I("name")), d)
    assert not f(E("==", I("age"), V("30")), d) and f(E("!=", I("age"), V("30")), d)  # kinds differ: unequal, no error
    assert f(E("==", I("missing"), V(None)), d) and f(E("==", I("n"), V(None)), d)
    assert f(E(">", I("name"), V("alice")), d) and not f(E(">", I("name"), V(3.0)), d)
    assert not f(I("age"), d) and f(I("ok"), d)                          # the result must be a bool
    assert not f(E("==", I("age"), V(30.0)), b"not json") and not f(E("==", I("age"), V(30.0)), b"[1]")
    assert f(F("EXISTS", I("missing")), d) and not f(F("EXISTS", I("x")), b"[1]")  # EXISTS = "no error"
    assert f(F("DOES_NOT_EXIST", I("missing")), d) and not f(F("DOES_NOT_EXIST", I("n")), d)
    assert not f(E("CONTAINS", I("age"), V("3")), d)
    assert f(E(">", DOT("tags", "length"), V(1.0)), d) and not f(E("==", DOT("n", "x"), V(None)), d)


def test_lowering_and_column_extraction():
    cols = {"age": 0, "status": 1, "user.email": 2}
    prog = hf.lower(E("AND", E(">=", I("age"), V(18.0)), E("ENDS_WITH", DOT("user", "email"), V("@x.org"))), cols)
    assert [o["op"] for o in prog] == [_capi.FOP_COL, _capi.FOP_NUM, _capi.FOP_GE, _capi.FOP_COL, _capi.FOP_ENDS_WITH, _capi.FOP_AND]
    assert prog[3]["arg"] == 2 and prog[4]["str"] == "@x.org"
    for bad in (E("==", I("age"), I("status")), E("==", I("unknown"), V(1.0)), E(">=", F("LENGTH", I("age")), V(3.0)),
                E("CONTAINS", I("status"), I("age")), ("any", I("age"), I("age"))):
        with pytest.raises(hf.Unsupported):
            hf.lower(bad, cols)
    kind, vals = hf.column_values(b'{"age": 3, "status": null, "user": {"email": "a@b"}}', ["age", "status", "user.email", "nope", "user.zip"])
    assert kind == hf.DOC_OBJECT and vals == [(hf.MV_NUMBER, 3.0), (hf.MV_NULL, None), (hf.MV_STRING, "a@b"), (hf.MV_MISSING, None), (hf.MV_ERROR, None)]
    assert hf.column_values(b"{", ["age"])[0] == hf.DOC_INVALID and hf.column_values(b"3", ["age"])[0] == hf.DOC_OTHER


def _stack_effect(prog):
    """The value-stack discipline szg_filter_mask checks before it runs a program."""
    sp = 0
    for o in prog:
        op = o["op"]
        if op in (_capi.FOP_COL, _capi.FOP_NUM, _capi.FOP_STR, _capi.FOP_BOOL, _capi.FOP_NULL, _capi.FOP_EXISTS, _capi.FOP_NOT_EXISTS):
            pops = 0
        elif op in (_capi.FOP_NOT, _capi.FOP_CONTAINS, _capi.FOP_STARTS_WITH, _capi.FOP_ENDS_WITH, _capi.FOP_STR_TABLE):
            pops = 1
        elif op in (_capi.FOP_IN, _capi.FOP_NOT_IN):
            pops = o["arg"] + 1
        else:
            pops = 2
        assert sp >= pops, (o, sp)
        sp += 1 - pops
        assert sp <= 16
    return sp


def test_every_covered_reference_case_lowers_to_a_well_formed_program():
    lowered = 0
    for name, tree, doc, want in REFERENCE_TREE_CASES + REFERENCE_QUERY_CASES:
        fields = sorted(set(_paths(tree)))
        cols = {f: i for i, f in enumerate(fields)}
        try:
            prog = hf.lower(tree, cols, dictionary=lambda: ["john_doe123", "x"])
        except hf.Unsupported:
            assert "LENGTH" in name, name  # the only reference case outside the covered subset
            continue
        assert _stack_effect(prog) == 1, name
        assert all(o.get("arg", 0) < 32 for o in prog if o["op"] in (_capi.FOP_COL, _capi.FOP_EXISTS, _capi.FOP_NOT_EXISTS))
        lowered += 1
    assert lowered == len(REFERENCE_TREE_CASES) + len(REFERENCE_QUERY_CASES) - 1


# ------------------------------------------------------------------------------------------------ device path
FIELDS = ["age", "status", "flag", "tags", "user.email", "name", "score", "tags.length"]
COLS = {f: i for i, f in enumerate(FIELDS)}
WORDS = ["active", "pending", "inactive", "Active", "", "zeta", "alpha", "alp", "report_7", "a@example.com", "b@example.org", "ünï"]


def _random_doc(rng):
    r = rng.random()
    if r < 0.02:
        return b"{not json"
    if r < 0.04:
        return rng.choice([b"[1, 2]", b"3", b'"str"', b"null"])
    d = {}
    if rng.random() < 0.85:
        d["age"] = rng.choice([rng.randint(0, 60), rng.randint(0, 60) + 0.5, None, "30", -0.0, 1e300])
    if rng.random() < 0.8:
        d["status"] = rng.choice(WORDS + [7, None, True])
    if rng.random() < 0.7:
        d["flag"] = rng.choice([True, False, None, 1, "true"])
    if rng.random() < 0.5:
        d["tags"] = rng.choice([[], ["x"], ["x", "y", "z"], {"k": 1}])
    if rng.random() < 0.6:
        d["user"] = rng.choice([{"email": rng.choice(WORDS)}, {"mail": "q"}, {}, None, 5, [1]])
    if rng.random() < 0.7:
        d["name"] = rng.choice(WORDS)
    if rng.random() < 0.7:
        d["score"] = rng.random() * 10
    return json.dumps(d).encode()


def _random_tree(rng, depth=0):
    field = lambda: rng.choice([I("age"), I("status"), I("flag"), I("tags"), DOT("user", "email"), I("name"), I("score"),
                                DOT("tags", "length")])
    lit = lambda: V(rng.choice([30.0, 0.0, 12.5, 3.0, "active", "alp", "zzz", "", True, False, None, float(rng.randint(0, 60))]))
    r = rng.random()
    if depth < 3 and r < 0.35:
        return E(rng.choice(["AND", "OR"]), _random_tree(rng, depth + 1), _random_tree(rng, depth + 1))
    if depth < 3 and r < 0.42:
        return E("NOT", None, _random_tree(rng, depth + 1))
    r = rng.random()
    if r < 0.45:
        a, b = field(), lit()
        if rng.random() < 0.2:
            a, b = b, a
        return E(rng.choice(["==", "!=", "<", "<=", ">", ">="]), a, b)
    if r < 0.6:
        return E(rng.choice(["IN", "NOT_IN"]), field(), A(*[lit() for _ in range(rng.randint(0, 4))]))
    if r < 0.8:
        return E(rng.choice(["CONTAINS", "STARTS_WITH", "ENDS_WITH", "MATCHES"]), field(), V(rng.choice(["a", "act", "@example", "e", "", "^a", "ve$"])))
    if r < 0.9:
        return F(rng.choice(["EXISTS", "DOES_NOT_EXIST"]), rng.choice([I("age"), I("status"), I("tags"), I("name")]))
    return rng.choice([field(), lit()])  # not a boolean expression: never passes unless the value is a bool


def _load(ix, docs):
    import syzgydb_b200 as szg  # noqa: F401
    n = len(docs)
    ids = np.arange(n, dtype=np.uint64) * 5 + 2
    rows = np.random.default_rng(3).uniform(-1, 1, size=(n, ix.dim))
    ix.encode(rows, ids=ids, upsert=True)
    kinds, vals = zip(*[hf.column_values(d, FIELDS) for d in docs])
    ix.meta_upsert(ids, kinds, list(range(len(FIELDS))), vals)
    return ids


def _passing(ix, mask_id):
    gi, _, _ = ix.search_radius(np.zeros(ix.dim), 1e9, mask_id=mask_id)
    return set(int(i) for i in gi)


@pytest.mark.gpu
def test_device_filter_matches_the_oracle_document_by_document():
    import syzgydb_b200 as szg
    rng = random.Random(11)
    docs = [_random_doc(rng) for _ in range(3000)]
    with szg.Index(8, 8, szg.EUCLIDEAN) as ix:
        ids = _load(ix, docs)
        trees = [t for _, t, _, _ in REFERENCE_TREE_CASES[:13] + REFERENCE_QUERY_CASES[:1]] + [_random_tree(rng) for _ in range(150)]
        ran = 0
        for tree in trees:
            try:
                prog = hf.lower(tree, COLS, dictionary=ix.meta_dictionary)
            except hf.Unsupported:
                continue
            m = ix.filter_mask(prog)
            want = {int(i) for i, d in zip(ids, docs) if of.filter_document(tree, d)}
            got = _passing(ix, m)
            assert got == want, (tree, sorted(got ^ want)[:5])
            ix.mask_destroy(m)
            ran += 1
        assert ran > 120


@pytest.mark.gpu
def test_device_filter_reference_cases_and_filtered_search():
    import syzgydb_b200 as szg
    from oracle import pyoracle as o
    from tests.common import assert_results_match
    # every golden case the lowering covers, each on its own one-document collection plus a decoy that must not pass
    for name, tree, doc, want in REFERENCE_TREE_CASES + REFERENCE_QUERY_CASES:
        fields = sorted({p for p in _paths(tree)})
        cols = {f: i for i, f in enumerate(fields)}
        with szg.Index(4, 64, szg.EUCLIDEAN) as ix:
            ids = np.array([7, 8], dtype=np.uint64)
            ix.encode(np.ones((2, 4)), ids=ids, upsert=True)
            try:
                prog = None
                kinds, vals = zip(*[hf.column_values(d, fields) for d in (doc.encode(), b"{}")])
                ix.meta_upsert(ids, kinds, list(range(len(fields))), vals)
                prog = hf.lower(tree, cols, dictionary=ix.meta_dictionary)
            except hf.Unsupported:
                assert "LENGTH" in name
                continue
            got = _passing(ix, ix.filter_mask(prog))
            assert (7 in got) is want, name
            assert (8 in got) == of.filter_document(tree, b"{}"), name
    # a filtered top-k through the device mask equals the oracle's filtered scan (bucket < 3)
    n, dims = 20000, 48
    codes = o.synth_rows(5, 0, n, dims, 8)
    ids = np.arange(n, dtype=np.uint64)
    docs = [json.dumps({"bucket": int(i % 10)}).encode() for i in range(n)]
    tree = E("<", I("bucket"), V(3.0))
    with szg.Index(dims, 8, szg.COSINE) as ix:
        ix.upsert(ids, codes)
        kinds, vals = zip(*[hf.column_values(d, ["bucket"]) for d in docs])
        ix.meta_upsert(ids, kinds, [0], vals)
        m = ix.filter_mask(hf.lower(tree, {"bucket": 0}))
        q = o.synth_queries(6, 0, 1, dims)[0]
        gi, gd, gn, scanned = ix.search_topk(q, 10, mask_id=m)
        passmask = np.array([of.filter_document(tree, d) for d in docs], dtype=np.uint8)
        ri, rd, _ = o.search_exact(codes, ids, dims, 8, szg.COSINE, q, k=10, passmask=passmask)
        assert scanned == n and passmask.sum() == 6000
        assert_results_match(gi[0, :gn[0]], gd[0, :gn[0]], ri, rd, what="filtered top-k")
        # a removed document's slot does not leak its metadata to the next document that takes it
        ix.remove(ids[:1])
        ix.upsert(np.array([999999], dtype=np.uint64), codes[:1])
        m2 = ix.filter_mask(hf.lower(tree, {"bucket": 0}))
        assert 999999 not in _passing(ix, m2) and 0 not in _passing(ix, m2) and 1 in _passing(ix, m2)


def _paths(node):
    if node is None:
        return
    p = hf._path(node)
    if p is not None:
        yield p
        return
    if node[0] == "expr":
        yield from _paths(node[2])
        yield from _paths(node[3])
    elif node[0] in ("array",):
        for e in node[1]:
            yield from _paths(e)
    elif node[0] == "func":
        for e in node[2]:
            yield from _paths(e)
