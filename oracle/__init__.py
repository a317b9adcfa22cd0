"""ORACLE — CPU restatement of the reference's hot path.  Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (score_based_multimodal_autoencoder_b200) never imports it."""
