"""Reference-facing model classes: `hulc_b200.models.hulc.Hulc` / `hulc_b200.models.gcbc.GCBC` answer to the Hydra
`_target_` roles of `hulc.models.hulc.Hulc` / `hulc.models.gcbc.GCBC` (conf/model/{hulc,mcil,gcbc}.yaml)."""
