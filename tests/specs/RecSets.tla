------------------------------- MODULE RecSets -------------------------------
(* Builder-authored regression spec: record sets with run-time components, membership in them, and
   "field" \in DOMAIN r on a tagged union (the constructs of AdvancedExamples/InnerSerial.tla:5-30). *)
EXTENDS Naturals, Sequences, FiniteSets
CONSTANT Proc
VARIABLES q, last
vars == <<q, last>>

BoundedSeq(S, n) == UNION {[1..k -> S] : k \in 0..n}
Op == [kind : {"rd"}, who : Proc] \cup [kind : {"wr"}, who : Proc, val : 1..2]
TypeOK == /\ q \in [Proc -> BoundedSeq(1..2, 2)]
          /\ last \in Op \cup {"none"}

Ids == UNION {[proc : {p}, idx : DOMAIN q[p]] : p \in Proc}        \* run-time record set

Init == q = [p \in Proc |-> << >>] /\ last = "none"
Push(p, v) == /\ Len(q[p]) < 2
              /\ q' = [q EXCEPT ![p] = Append(@, v)]
              /\ last' = [kind |-> "wr", who |-> p, val |-> v]
Peek(p) == /\ Len(q[p]) > 0
           /\ last' = [kind |-> "rd", who |-> p]
           /\ UNCHANGED q
Next == \E p \in Proc : (\E v \in 1..2 : Push(p, v)) \/ Peek(p)
Spec == Init /\ [][Next]_vars

IdsOK   == /\ Cardinality(Ids) = Len(q[CHOOSE p \in Proc : TRUE]) + Cardinality({i \in Ids : i.proc # (CHOOSE p \in Proc : TRUE)})
           /\ \A i \in Ids : i \in [proc : Proc, idx : DOMAIN q[i.proc]]
           /\ \A p \in Proc : [proc |-> p, idx |-> 3] \notin Ids
DomOK   == last = "none" \/ (("val" \in DOMAIN last) <=> (last.kind = "wr"))
MemOK   == last = "none" \/ last \in [kind : {"rd"}, who : Proc] \/ last \in [kind : {"wr"}, who : Proc, val : 1..2]
=============================================================================
