------------------------------- MODULE SubsetsWide -------------------------------
EXTENDS Naturals, FiniteSets
VARIABLES s, pick
vars == <<s, pick>>
U == 1..40
TypeOK == s \subseteq U /\ pick \subseteq U
Init == s = {} /\ pick = {}
Add == \E x \in {3, 31, 32, 33, 40} : x \notin s /\ s' = s \cup {x} /\ pick' \in SUBSET s' /\ pick \subseteq pick'
Next == Add
Spec == Init /\ [][Next]_vars
PickOK == pick \subseteq s /\ \A R \in SUBSET s : Cardinality(R) <= Cardinality(s)
=============================================================================
