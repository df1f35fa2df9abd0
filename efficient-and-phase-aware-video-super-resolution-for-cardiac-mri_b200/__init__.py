"""B200-native RefineNet hot path. Put this directory on sys.path to get `pvsr` (engine) and `src` (the
reference-facing mirror: src.main, src.model.nets.RefineNet, src.runner.*)."""
