-------------------------------- MODULE race --------------------------------
(* Builder-authored demo (not from the reference): the classic lost update -- two processes
   read a shared counter and write back the increment in separate steps. *)
EXTENDS Naturals

(* --algorithm race
variables counter = 0;

process Inc \in 1..2
  variables tmp = 0;
begin
rd: tmp := counter;
wr: counter := tmp + 1;
end process

end algorithm *)

Correct == (\A p \in 1..2 : pc[p] = "Done") => counter = 2
=============================================================================
