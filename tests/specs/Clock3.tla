------------------------------- MODULE Clock3 -------------------------------
(* Test model (builder-authored): a mod-3 counter with two refinement properties,
   one that holds (steps of +1 mod 3) and one that does not (claims steps of +2). *)
EXTENDS Naturals
VARIABLE c
Init == c = 0
Next == c' = (c + 1) % 3
Spec == Init /\ [][Next]_c
StepOne == c' = (c + 1) % 3
StepTwo == c' = (c + 2) % 3
Good == (c \in 0..2) /\ [][StepOne]_c
Bad == (c = 0) /\ [][StepTwo]_c
BadInit == (c = 1) /\ [][StepOne]_c
=============================================================================
