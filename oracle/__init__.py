"""CPU oracle (test infrastructure). See oracle/oracle.py for the rules on who may import it."""
