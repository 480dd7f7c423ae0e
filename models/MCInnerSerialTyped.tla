------------------------- MODULE MCInnerSerialTyped -------------------------
(* Builder-authored typing wrapper for AdvancedExamples/MCInnerSerial (found through the module search path): the
   spec's own DataInvariant types opQ and opOrder through the state-dependent set opId, which gives the compiler no
   fixed universe.  TypeOK below bounds them by the CONSTRAINT (queues of at most MaxQLen operations, one more in a
   generated-then-discarded successor).  It is used for typing only; the cfg is MCInnerSerial.cfg's. *)
EXTENDS MCInnerSerial

BoundedSeq(S, n) == UNION {[1..k -> S] : k \in 0..n}
OpIdU  == [proc : Proc, idx : 1..(MaxQLen + 1)]
OpValU ==      [req : Request,   reg : Reg]
          \cup [req : WrRequest, reg : {Done}]
          \cup [req : RdRequest, reg : {Done}, source : OpIdU \cup {InitWr}]

TypeOK == /\ regFile \in [Proc -> [Reg -> RegValue]]
          /\ opQ \in [Proc -> BoundedSeq(OpValU, MaxQLen + 1)]
          /\ opOrder \subseteq (OpIdU \X OpIdU)
=============================================================================
