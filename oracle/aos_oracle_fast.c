/* placeholder */ typedef int aos_oracle_fast_placeholder;
