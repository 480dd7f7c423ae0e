------------------------------ MODULE MCPaxos3 ------------------------------
(* Builder-authored model of BASELINE config #3: Lamport's Paxos spec
   (examples/Paxos/Paxos.tla in the reference, found through the module search
   path) with 3 acceptors, 2 values and ballots 0..MaxBallot -- the alternates
   left commented out at examples/Paxos/MCPaxos.tla:7-9. *)
EXTENDS Paxos, TLC
CONSTANTS a1, a2, a3, v1, v2, MaxBallot

MCAcceptor == {a1, a2, a3}
MCValue    == {v1, v2}
MCQuorum   == {{a1, a2}, {a1, a3}, {a2, a3}}
MCBallot   == 0..MaxBallot
MCSymmetry == Permutations(MCAcceptor) \cup Permutations(MCValue)   \* as in the reference's MCPaxos.tla:12

Inv1 == Inv!1
Inv2 == Inv!2
Inv3 == Inv!3
Inv4 == Inv!4
=============================================================================
