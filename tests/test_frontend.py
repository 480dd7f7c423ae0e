"""Front end: parser over the whole reference corpus, cfg grammar, PlusCal translation layout."""
import glob
import os

import pytest

from conftest import REF, needs_reference
from tla_rust_b200.front.parser import parse_module_text, parse_expr_text, read_text
from tla_rust_b200.front.spec import parse_cfg, Model
from tla_rust_b200.front.pcal import translate_text
from tla_rust_b200.front.values import ModelValue, fmt


@needs_reference
def test_parse_whole_corpus():
    # all 84 modules, including the Standard/ ones with instance-qualified infix operators (a R!+ b) and -. a == ...
    files = glob.glob(REF + "/**/*.tla", recursive=True)
    assert len(files) >= 84
    for f in files:
        parse_module_text(read_text(f))


def test_junction_lists_and_precedence():
    e = parse_expr_text("/\\ a = 1\n/\\ \\/ b\n   \\/ c\n/\\ d")
    assert e.k == "and" and len(e.a[0]) == 3 and e.a[0][1].k == "or"
    e = parse_expr_text("a + b * c - d")
    # TLA+ table: `-` (11-11) binds tighter than `+` (10-10), `*` (13-13) tighter than both
    assert e.a[0] == "+" and e.a[2].a[0] == "-" and e.a[2].a[1].a[0] == "*"
    e = parse_expr_text("[f EXCEPT ![a].b = @ + 1, ![c] = 2]")
    assert e.k == "except" and len(e.a[1]) == 2
    e = parse_expr_text("{x \\in S : x > 1} \\cup {f[x] : x \\in S}")
    assert e.a[1].k == "setfilter" and e.a[2].k == "setmap"
    e = parse_expr_text("Inv!2 /\\ V!ShowsSafeAt(Q, b, v) /\\ Thm!:")
    assert [x.k for x in e.a[0][0].a[0]] + [e.a[0][1].k] == ["sel", "sel", "sel"] or e.k == "and"


def test_cfg_grammar():
    # TLC/ConfigFileGrammar.tla:8-33 + MCPaxos.cfg:9 module-scoped override + comments
    c = parse_cfg("""SPECIFICATION Spec \\* c1
    CONSTANTS a1=a1 Acceptor <- MCAcceptor N = 3 S = {x, "s", 2} (* c2 *)
      Ballot <-[Voting] MCBallot
    INVARIANT Inv1 Inv2
    PROPERTY P SYMMETRY Sym CONSTRAINT C ACTION-CONSTRAINT AC""")
    assert c.specification == "Spec" and c.invariants == ["Inv1", "Inv2"]
    assert ("a1", ModelValue("a1")) in c.const_assign and ("N", 3) in c.const_assign
    assert ("S", frozenset({ModelValue("x"), "s", 2})) in c.const_assign
    assert ("Ballot", "Voting", "MCBallot") in c.const_subst and ("Acceptor", None, "MCAcceptor") in c.const_subst
    assert c.symmetry == "Sym" and c.constraints == ["C"] and c.action_constraints == ["AC"] and c.properties == ["P"]


README_BUGGY = (("     alice_account := alice_account - money;", "     A: alice_account := alice_account - money;"),
                ("     bob_account := bob_account + money;", "     B: bob_account := bob_account + money;"))


@needs_reference
def test_pcal_layout_matches_readme_locations():
    """The README trace names Transfer(self) as 'line 35, col 19 to line 40, col 42' etc (README.md:278-306);
    our translator must put the actions on exactly those lines/columns."""
    src = open(REF + "/pcal_intro.tla").read()
    for a, b in README_BUGGY:
        src = src.replace(a, b)
    out, had = translate_text(src)
    assert had
    lines = out.split("\n")
    assert lines[34].startswith("Transfer(self) == /\\ pc[self] = \"Transfer\"")
    assert lines[39] == " " * 34 + "money >>" and len(lines[39]) == 42
    assert lines[41].startswith("A(self) == ") and len(lines[44]) == 63
    assert lines[46].startswith("B(self) == ") and len(lines[49]) == 65
    assert lines[52].rstrip().endswith("Assert(alice_account >= 0,") and len(lines[53]) == 66
    assert '"Failure of assertion at line 16, column 4."' in lines[53]


@needs_reference
def test_assumes_and_printvalues():
    m = Model(REF + "/examples/SpecifyingSystems/SimpleMath/SimpleMath.tla")
    assert all(v is True for _, v in m.check_assumes())
    m = Model(REF + "/examples/SpecifyingSystems/AsynchronousInterface/PrintValues.tla")
    import io
    m.ev.out = io.StringIO()
    m.check_assumes()
    assert m.ev.print_out[0] == '<<"Three more cats: ", 4>>'
    m = Model(REF + "/examples/Paxos/MCVoting.tla")
    assert len(m.check_assumes()) == 2
