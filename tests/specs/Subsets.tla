------------------------------- MODULE Subsets -------------------------------
(* Builder-authored regression spec: SUBSET of a run-time set as an assignment domain, as a quantifier domain and
   under a set filter (AdvancedExamples/InnerSerial.tla:47-67 `totalOpOrder`, `UpdateOpOrder`). *)
EXTENDS Naturals, FiniteSets
CONSTANT Elem
VARIABLES s, pick
vars == <<s, pick>>

TypeOK == s \subseteq Elem /\ pick \subseteq Elem

Init == s = {} /\ pick = {}
Add(x) == /\ x \notin s
          /\ s' = s \cup {x}
          /\ pick' \in SUBSET s'
          /\ pick \subseteq pick'
Shrink == /\ pick' \in {p \in SUBSET pick : Cardinality(p) + 1 = Cardinality(pick)}
          /\ UNCHANGED s
Next == (\E x \in Elem : Add(x)) \/ Shrink
Spec == Init /\ [][Next]_vars

Refl  == {R \in SUBSET (s \X s) : \A a \in s : <<a, a>> \in R}
ReflOK == /\ \E R \in Refl : \A a, b \in s : (a # b) => <<a, b>> \notin R          \* the identity relation is in Refl
          /\ \A R \in Refl : Cardinality(R) >= Cardinality(s)
PickOK == pick \in SUBSET s
=============================================================================
