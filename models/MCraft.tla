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

(* Typing of the bounded model.  raft.tla has no type invariant; the compiler needs one to lay the state out in a
   fixed number of bits.  A successor may exceed the CONSTRAINT by one step (it is generated, then discarded),
   hence the "+ 1" bounds.  TypeOK is a true invariant of the constrained model (checked in
   tests/test_raft_device.py) but is not listed in the cfg: it is only used for typing. *)
BoundedSeq(S, n)  == UNION {[1..k -> S] : k \in 0..n}
PartialFcn(D, R)  == UNION {[d -> R] : d \in SUBSET D}
Terms    == 1..(MaxTerm + 1)
LogTerms == 0..(MaxTerm + 1)
LogIdx   == 0..(MaxLogLen + 1)
Entry    == [term : Terms, value : 1..MaxClientRequests]
Logs     == BoundedSeq(Entry, MaxLogLen + 1)
Message  ==
         [mtype : {RequestVoteRequest}, mterm : Terms, mlastLogTerm : LogTerms, mlastLogIndex : LogIdx,
          msource : Server, mdest : Server]
    \cup [mtype : {RequestVoteResponse}, mterm : Terms, mvoteGranted : BOOLEAN, mlog : Logs,
          msource : Server, mdest : Server]
    \cup [mtype : {AppendEntriesRequest}, mterm : Terms, mprevLogIndex : LogIdx, mprevLogTerm : LogTerms,
          mentries : BoundedSeq(Entry, 1), mlog : Logs, mcommitIndex : LogIdx, msource : Server, mdest : Server]
    \cup [mtype : {AppendEntriesResponse}, mterm : Terms, msuccess : BOOLEAN, mmatchIndex : LogIdx,
          msource : Server, mdest : Server]

TypeOK ==
    /\ DOMAIN messages \subseteq Message
    /\ \A m \in DOMAIN messages : messages[m] \in 0..2
    /\ Cardinality(DOMAIN messages) <= MaxMessages + 1
    /\ elections \subseteq [eterm : Terms, eleader : Server, elog : Logs, evotes : SUBSET Server,
                             evoterLog : PartialFcn(Server, Logs)]
    /\ allLogs \subseteq Logs
    /\ currentTerm \in [Server -> Terms]
    /\ state \in [Server -> {Follower, Candidate, Leader}]
    /\ votedFor \in [Server -> Server \cup {Nil}]
    /\ clientRequests \in 1..MaxClientRequests
    /\ log \in [Server -> Logs]
    /\ commitIndex \in [Server -> LogIdx]
    /\ committedLog \in Logs
    /\ committedLogDecrease \in BOOLEAN
    /\ votesSent \in [Server -> BOOLEAN]
    /\ votesGranted \in [Server -> SUBSET Server]
    /\ voterLog \in [Server -> PartialFcn(Server, Logs)]
    /\ nextIndex \in [Server -> [Server -> 1..(MaxLogLen + 2)]]
    /\ matchIndex \in [Server -> [Server -> LogIdx]]
=============================================================================
