-------------------------------- MODULE lock --------------------------------
(* Builder-authored demo (not from the reference): Peterson's mutual exclusion for two
   processes, written in PlusCal p-syntax like the reference's pcal_intro.tla. *)
EXTENDS Naturals

(* --algorithm peterson
variables flag = [i \in {0, 1} |-> FALSE], turn = 0;

process Proc \in {0, 1}
begin
a1: flag[self] := TRUE;
a2: turn := 1 - self;
a3: await flag[1 - self] = FALSE \/ turn = self;
cs: skip;
a4: flag[self] := FALSE;
    goto a1;
end process

end algorithm *)

MutualExclusion == ~(pc[0] = "cs" /\ pc[1] = "cs")
=============================================================================
