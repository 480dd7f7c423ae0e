----------------------------- MODULE Containers -----------------------------
(* Builder-authored regression spec for the device lowering of dynamically-shaped values: a message bag
   (function with a dynamic domain), a set of sequences, a partial function, bounded sequences with
   SubSeq / SelectSeq, a run-time cartesian product, UNION over a computed family, CHOOSE with a tuple
   pattern and a RECURSIVE operator.  The constructs are the ones raft.tla and
   serializableSnapshotIsolation.tla of the reference use; the oracle (AST evaluator) defines the expected
   counts (tests/test_containers.py). *)
EXTENDS Naturals, Sequences, FiniteSets, TLC
CONSTANTS Node, Max
VARIABLES bag, seen, pf, q, edges
vars == <<bag, seen, pf, q, edges>>

BoundedSeq(S, n) == UNION {[1..k -> S] : k \in 0..n}
PartialFcn(D, R) == UNION {[d -> R] : d \in SUBSET D}
Vals == 1..Max
Msg  == [src : Node, body : BoundedSeq(Vals, 2)]

TypeOK == /\ DOMAIN bag \subseteq Msg
          /\ \A m \in DOMAIN bag : bag[m] \in 0..2
          /\ Cardinality(DOMAIN bag) <= 3
          /\ seen \subseteq BoundedSeq(Vals, 2)
          /\ pf \in PartialFcn(Node, 1..3)
          /\ q \in BoundedSeq(Vals, 3)
          /\ edges \subseteq Node \X Node

Range(f) == {f[x] : x \in DOMAIN f}
Live == {m \in DOMAIN bag : bag[m] > 0}

RECURSIVE Reach(_, _)
Reach(n, vis) ==
    IF n \in vis THEN vis
    ELSE LET next == {e[2] : e \in {e \in edges : e[1] = n}}
         IN  (vis \cup {n}) \cup UNION {Reach(m, vis \cup {n}) : m \in next}

Init == /\ bag = [m \in {} |-> 0]
        /\ seen = {}
        /\ pf = [n \in {} |-> 1]
        /\ q = << >>
        /\ edges = {}

Send(n) ==
    LET m == [src |-> n, body |-> SubSeq(q, 1, IF Len(q) < 2 THEN Len(q) ELSE 2)]
    IN  /\ Cardinality(DOMAIN bag) < 2 \/ m \in DOMAIN bag
        /\ bag' = IF m \in DOMAIN bag THEN [bag EXCEPT ![m] = IF @ < 2 THEN @ + 1 ELSE 2]
                                      ELSE bag @@ (m :> 1)
        /\ UNCHANGED <<seen, pf, q, edges>>

Recv(m) ==
    /\ bag' = [bag EXCEPT ![m] = @ - 1]
    /\ seen' = seen \cup {m.body}
    /\ pf' = pf @@ (m.src :> Len(m.body) + 1)
    /\ UNCHANGED <<q, edges>>

Push(v) == /\ Len(q) < 2
           /\ q' = Append(q, v)
           /\ UNCHANGED <<bag, seen, pf, edges>>

Drop == /\ q' = SelectSeq(q, LAMBDA x : x > 1)
        /\ q' /= q
        /\ UNCHANGED <<bag, seen, pf, edges>>

Link(a, b) == /\ Cardinality(edges) < 2 /\ a /= b
              /\ <<a, b>> \notin edges
              /\ edges' = edges \cup {<<a, b>>}
              /\ UNCHANGED <<bag, seen, pf, q>>

Next == \/ \E n \in Node : Send(n)
        \/ \E m \in Live : Recv(m)
        \/ \E v \in Vals : Push(v)
        \/ Drop
        \/ \E a, b \in Node : Link(a, b)

Spec == Init /\ [][Next]_vars

ReachOK  == \A n \in Node : n \in Reach(n, {}) /\ Reach(n, {}) \subseteq Node
PfOK     == \A n \in DOMAIN pf : pf[n] \in 1..3
ProdOK   == edges \subseteq {e[1] : e \in edges} \X {e[2] : e \in edges}
PickOK   == edges = {} \/ (LET e == CHOOSE <<a, b>> \in edges : TRUE IN e \in edges)
UnionOK  == UNION Range([n \in Node |-> IF n \in DOMAIN pf THEN {pf[n]} ELSE {}]) \subseteq 1..3
SeenOK   == \A s \in seen : Len(s) <= 2
BagOK    == \A m \in DOMAIN bag : bag[m] \in 0..2
=============================================================================
