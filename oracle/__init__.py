"""Oracle (test infrastructure): CPU restatement of the reference UAHN forward + fixtures tooling."""
