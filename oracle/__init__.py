"""oracle/ — CPU restatement + reference harness. TEST INFRASTRUCTURE ONLY (never imported by tampc_b200)."""
