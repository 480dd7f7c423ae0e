------------------------------ MODULE mac ------------------------------
EXTENDS Naturals, TLC
(* --algorithm mac
variables x = 0, y = 0;
macro bump(v, by) begin
  v := v + by;
end macro;
macro both(k) begin
  bump(x, k);
  if x > 2 then y := y + 1; end if;
end macro;
process P \in 1..2
begin
  A: bump(x, 1);
  B: both(2);
  C: assert x <= 6;
end process
end algorithm *)
=============================================================================
