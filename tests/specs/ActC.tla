------------------------------ MODULE ActC ------------------------------
(* Builder-authored test spec for ACTION-CONSTRAINT (TLC/ConfigFileGrammar.tla:8-12): a constraint on the pair
   <<state, successor>>.  Steps that break it are generated and counted, their target is neither stored nor explored:
   with SmallStep the jumps are cut and 6..8 stay unreachable. *)
EXTENDS Naturals
VARIABLES x, y
Init == x = 0 /\ y = 0
Next == \/ x < 5 /\ x' = x + 1 /\ y' = y
        \/ x + 3 <= 8 /\ x' = x + 3 /\ y' = y
        \/ y' = (y + 1) % 3 /\ x' = x
SmallStep == x' - x < 3
TypeOK == x \in 0..8 /\ y \in 0..2
==========================================================================
