"""CPU oracle (test infrastructure only). See oracle/orb_oracle.h. Import `oracle.binding`."""
