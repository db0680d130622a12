# The importable package is umi-collapse-rs_b200/umigpu (a hyphenated directory cannot be imported by name).
