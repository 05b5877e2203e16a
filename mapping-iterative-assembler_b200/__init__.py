"""miagpu: B200-native hot path for the Mapping Iterative Assembler (MIA).

Import as ``mia_b200`` via the repo-root helper ``_pkg.load()``.
"""
