------------------------------ MODULE Cinit ------------------------------
(* Builder-authored test spec: some initial states lie outside the CONSTRAINT.  TLC generates and counts them but
   never explores them (ConfigFileGrammar.tla:8-12; FIFO/MCInnerFIFO.cfg:23-31 for the meaning of CONSTRAINT). *)
EXTENDS Naturals
VARIABLE x
Init == x \in 0..5
Next == x' = (x + 2) % 9
Small == x < 3
TypeOK == x \in 0..8
==========================================================================
