------------------------------- MODULE MCraft -------------------------------
(* Builder-authored model of BASELINE config #4: Ongaro's Raft spec as modified in the reference
   (examples/raft.tla, found through the module search path).  The reference ships neither a .cfg nor an
   MC module for it and defines no MaxTerm / MaxLogLen: the state space is bounded here by a CONSTRAINT
   (SURVEY.md §7 step 1).  Server ids and the message-type / state constants (raft.tla:11-24) are model
   values.  The candidate invariant is the negation of MoreThanOneLeader (raft.tla:506-507). *)
EXTENDS raft

CONSTANTS MaxTerm, MaxLogLen, MaxMessages

StateConstraint ==
    /\ \A i \in Server : currentTerm[i] <= MaxTerm
    /\ \A i \in Server : Len(log[i]) <= MaxLogLen
    /\ Cardinality(DOMAIN messages) <= MaxMessages

AtMostOneLeaderPerTerm == ~MoreThanOneLeader
=============================================================================
