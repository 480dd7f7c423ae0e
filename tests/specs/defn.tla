------------------------------ MODULE defn ------------------------------
EXTENDS Naturals, TLC
(* --algorithm defn
variables x = 0, y = 0;
define
  Sum == x + y
  Small(n) == n < 3
end define;
macro bump(v, by) begin
  v := v + by;
end macro;
process P \in 1..2
begin
  A: await Small(Sum);
     bump(x, 1);
  B: bump(y, self);
  C: assert Sum <= 7;
end process
end algorithm *)
=============================================================================
