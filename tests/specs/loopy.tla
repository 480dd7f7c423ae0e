------------------------------ MODULE loopy ------------------------------
EXTENDS Naturals, TLC
(* --algorithm loopy
variables total = 0, flag = FALSE;
process Worker \in 1..2
variables i = 0, got = 0;
begin
  L1: while i < 2 do
        either
          total := total + 1;
        or
          with d \in {1, 2} do
            got := d;
          end with;
        end either;
        i := i + 1;
      end while;
  L2: if got = 2 then
        flag := TRUE;
      else
        skip;
      end if;
  L3: await total >= 0;
      assert total <= 4;
end process
end algorithm *)
Bound == total <= 4
=============================================================================
