"""CPU oracle for the Ackermann env-step hot path (TEST INFRASTRUCTURE, never on the product path)."""
