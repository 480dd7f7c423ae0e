------------------------------- MODULE MCssi -------------------------------
(* Builder-authored model of BASELINE config #5: Cahill's serializable snapshot isolation
   (examples/serializableSnapshotIsolation.tla in the reference, found through the module search
   path).  The reference ships no .cfg for it (its intended model is described in prose at
   serializableSnapshotIsolation.tla:26-96): TxnId and Key are sets of model values, NoLock must
   be bound to a model value because its definition (line 24) is an unbounded CHOOSE, deadlock
   checking stays ON (line 57), and the invariants are the "should never be violated" list
   (lines 59-79). *)
EXTENDS serializableSnapshotIsolation

(* cfg INVARIANTs must be identifiers: wrap the parameterised predicates of lines 59-79 *)
InvWellFormed == WellFormedTransactionsInHistory(history)
InvCahill     == CahillSerializable(history)
InvBernstein  == BernsteinSerializable(history)
=============================================================================
