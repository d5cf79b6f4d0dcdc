"""Known-answer vectors ported from the reference's own unit tests:
crates/shared/src/models/node.rs:659-1241 (33 tests).  Each MEETS row is
(test name, reference line, specs kwargs, requirement string, expected).
Specs kwargs follow create_compute_specs (node.rs:626-657): gpu is Some iff any
gpu field is given; cpu is Some iff cores is given."""

S = dict  # specs helper


def specs(gpu_count=None, gpu_model=None, gpu_mem=None, cpu_cores=None, ram=None, storage=None, **kw):
    d = dict(gpu_count=gpu_count, gpu_model=gpu_model, gpu_mem=gpu_mem, cpu_cores=cpu_cores, ram=ram, storage=storage)
    d.update(kw)
    return d


REQ_FULL = "gpu:count=4;gpu:model=A100;gpu:memory_mb=40000;cpu:cores=16;ram_mb=64000;storage_gb=500"
REQ_OR = ("gpu:count=8;gpu:model=H100;gpu:memory_mb=80000 ; gpu:count=16;gpu:model=A100;gpu:memory_mb=80000 ; "
          "ram_mb=128000; storage_gb=1000")
A100_SPEC_4x40 = specs(4, "A100", 40000, 16, 64000, 500)

MEETS = [
    ("test_meets_exact_match", 741, specs(4, "nvidia_a100_80gb_pcie", 40000, 16, 64000, 500), REQ_FULL, True),
    ("test_a100_range_case", 755, specs(1, "nvidia_a100_80gb_pcie", 40000, 16, 64000, 700),
     "gpu:count=4;gpu:model=a100,h100,h200;gpu:count=1;gpu:model=a100,h100,h200;storage_gb=700", True),
    ("test_meets_more_than_required", 770, specs(8, "NVIDIA A100 80GB", 80000, 32, 128000, 1000),
     "gpu:count=8;gpu:model=A100;gpu:memory_mb=40000;cpu:cores=16;ram_mb=64000;storage_gb=500", True),
    ("test_meets_fails_ram", 786, specs(4, "A100", 40000, 16, 32000, 500), REQ_FULL, False),
    ("test_meets_fails_gpu_count", 801, specs(2, "A100", 40000, 16, 64000, 500), REQ_FULL, False),
    ("test_meets_fails_gpu_model", 816, specs(4, "RTX 3090", 24000, 16, 64000, 500), REQ_FULL, False),
    ("test_meets_gpu_or_option1", 831, specs(8, "NVIDIA H100", 80000, 64, 256000, 2000), REQ_OR, True),
    ("test_meets_gpu_or_option2", 848, specs(16, "NVIDIA A100", 80000, 64, 256000, 2000), REQ_OR, True),
    ("test_meets_gpu_or_fails_both", 865, specs(4, "NVIDIA A100", 80000, 64, 256000, 2000), REQ_OR, False),
    ("test_meets_no_gpu_required/with_gpu", 882, specs(1, "RTX 3060", 12000, 8, 32000, 500),
     "ram_mb=16000;storage_gb=200;cpu:cores=4", True),
    ("test_meets_no_gpu_required/no_gpu", 882, specs(None, None, None, 8, 32000, 500),
     "ram_mb=16000;storage_gb=200;cpu:cores=4", True),
    ("test_meets_gpu_required_node_has_none", 903, specs(None, None, None, 8, 32000, 500),
     "gpu:count=1;gpu:model=A100;ram_mb=16000", False),
    ("test_meets_optional_fields_in_req", 913, specs(4, "NVIDIA H100", 80000, 64, 256000, 2000),
     "gpu:count=4; ram_mb=128000", True),
    ("test_meets_optional_fields_in_spec/mem", 933, specs(4, "A100", None, 16, 64000, 500),
     "gpu:count=4;gpu:model=A100;gpu:memory_mb=40000", False),
    ("test_meets_optional_fields_in_spec/no_mem", 933, specs(4, "A100", None, 16, 64000, 500),
     "gpu:count=4;gpu:model=A100", True),
    ("test_meets_min_max_gpu_memory/in_range", 960, A100_SPEC_4x40,
     "gpu:count=4;gpu:model=A100;gpu:memory_mb_min=30000;gpu:memory_mb_max=50000", True),
    ("test_meets_min_max_gpu_memory/below", 960, A100_SPEC_4x40,
     "gpu:count=4;gpu:model=A100;gpu:memory_mb_min=20000;gpu:memory_mb_max=35000", False),
    ("test_meets_min_max_gpu_memory/above", 960, A100_SPEC_4x40,
     "gpu:count=4;gpu:model=A100;gpu:memory_mb_min=45000;gpu:memory_mb_max=60000", False),
    ("test_gpu_model_case_insensitive_matching", 1015, specs(1, "NVIDIA_A100_80GB_PCIE", 40000),
     "gpu:model=nvidia_a100_80gb_pcie", True),
    ("test_gpu_model_no_match", 1030, specs(1, "AMD Radeon RX 7900", 20000), "gpu:model=nvidia,rtx", False),
    ("test_complex_gpu_or_logic", 1066, specs(2, "rtx4090", 24000, None, 64000, None),
     "gpu:count=8;gpu:model=H100;gpu:memory_mb=80000;gpu:count=4;gpu:model=A100;gpu:memory_mb=40000;"
     "gpu:count=2;gpu:model=RTX4090;gpu:memory_mb=24000;ram_mb=62000", True),
    ("test_meets_total_memory_requirements/1", 1119, specs(4, "NVIDIA A100", 40000),
     "gpu:count=4;gpu:model=A100;gpu:total_memory_min=120000;gpu:total_memory_max=200000", True),
    ("test_meets_total_memory_requirements/2", 1119, specs(4, "NVIDIA A100", 40000),
     "gpu:count=4;gpu:model=A100;gpu:total_memory_min=200000", False),
    ("test_meets_total_memory_requirements/3", 1119, specs(4, "NVIDIA A100", 40000),
     "gpu:count=4;gpu:model=A100;gpu:total_memory_max=120000", False),
    ("test_meets_total_memory_requirements/4", 1119, specs(4, "NVIDIA A100", 40000),
     "gpu:count=4;gpu:model=A100;gpu:total_memory_min=160000;gpu:total_memory_max=160000", True),
    ("test_meets_total_memory_missing_fields/no_memory", 1160, specs(4, "A100", None),
     "gpu:model=A100;gpu:total_memory_min=120000", True),
    ("test_meets_total_memory_missing_fields/no_count", 1160, specs(None, "A100", 40000),
     "gpu:model=A100;gpu:total_memory_min=120000", True),
    ("test_meets_total_memory_or_logic", 1198, specs(8, "NVIDIA H100", 80000),
     "gpu:count=4;gpu:model=A100;gpu:total_memory_min=160000;gpu:count=8;gpu:model=H100;gpu:total_memory_min=500000",
     True),
    ("test_complex_total_memory_scenario", 1215, specs(2, "RTX 4090", 24000),
     "gpu:count=8;gpu:model=H100;gpu:total_memory_min=600000;gpu:count=4;gpu:model=A100;"
     "gpu:total_memory_min=160000;gpu:count=2;gpu:model=RTX4090;gpu:total_memory_min=40000;"
     "gpu:total_memory_max=60000", True),
    # allocator-level vectors from node_groups/tests.rs:303-506 (memory_mb: Some(24))
    ("ng_requirements_rtx4090", 303, specs(8, "RTX4090", 24), "gpu:count=8;gpu:model=RTX4090;", True),
]

# (test name, line, string, expect_ok)
PARSE_OK = [
    ("test_requirements_parser_invalid/abc", 733, "gpu:count=abc", False),
    ("test_requirements_parser_invalid/key", 734, "ram_mb=100;gpu_model=xyz", False),
    ("test_requirements_parser_invalid/format", 735, "gpu:count=1=2", False),
    ("test_meets_min_max_gpu_memory/both", 1006,
     "gpu:count=4;gpu:model=A100;gpu:memory_mb=40000;gpu:memory_mb_min=30000;gpu:memory_mb_max=50000", False),
    ("test_parser_empty_values/count", 1046, "gpu:count=", False),
    ("test_parser_empty_values/model", 1047, "gpu:model=", True),
    ("test_gpu_memory_min_max_validation/bad", 1085, "gpu:memory_mb_min=40000;gpu:memory_mb_max=20000", False),
    ("test_gpu_memory_min_max_validation/ok", 1089, "gpu:memory_mb_min=20000;gpu:memory_mb_max=40000", True),
    ("test_total_memory_validation/bad", 1110, "gpu:total_memory_min=400000;gpu:total_memory_max=200000", False),
    ("test_total_memory_validation/ok", 1114, "gpu:total_memory_min=200000;gpu:total_memory_max=400000", True),
    ("test_parser_empty_string", 1059, "", True),
    ("extra/missing_equals", 195, "gpu:count", False),
    ("extra/plus_sign", 213, "gpu:count=+4", True),
    ("extra/overflow", 213, "gpu:count=4294967296", False),
    ("extra/max_before_min_bad", 274, "gpu:memory_mb_max=100;gpu:memory_mb_min=200", False),
]

# (test name, line, string, expected structure): gpu = list of dicts of the Some fields
PARSE_STRUCT = [
    ("test_requirements_parser_simple", 660,
     "gpu:count=1;gpu:model=A100;gpu:memory_mb=40000;ram_mb=64000;storage_gb=500",
     dict(gpu=[dict(count=1, model="A100", memory_mb=40000)], ram_mb=64000, storage_gb=500, cpu=None)),
    ("test_requirements_parser_gpu_or_logic", 675, REQ_OR,
     dict(gpu=[dict(count=8, model="H100", memory_mb=80000), dict(count=16, model="A100", memory_mb=80000)],
          ram_mb=128000, storage_gb=1000, cpu=None)),
    ("test_requirements_parser_gpu_minimal", 702, "gpu:count=8;gpu:model=H100 ; gpu:count=16; ram_mb=128000",
     dict(gpu=[dict(count=8, model="H100"), dict(count=16)], ram_mb=128000, storage_gb=None, cpu=None)),
    ("test_requirements_parser_no_gpu", 720, "ram_mb=32000;storage_gb=250;cpu:cores=8",
     dict(gpu=[], ram_mb=32000, storage_gb=250, cpu=8)),
    ("test_parser_whitespace_handling", 1051, " gpu:count = 1 ; gpu:model = A100 ; ram_mb = 32000 ",
     dict(gpu=[dict(count=1, model="A100")], ram_mb=32000, storage_gb=None, cpu=None)),
    ("test_parser_empty_string", 1059, "", dict(gpu=[], ram_mb=None, storage_gb=None, cpu=None)),
    ("test_total_memory_parsing", 1094,
     "gpu:count=4;gpu:model=A100;gpu:total_memory_min=160000;gpu:total_memory_max=320000",
     dict(gpu=[dict(count=4, model="A100", total_memory_min=160000, total_memory_max=320000)],
          ram_mb=None, storage_gb=None, cpu=None)),
    ("test_multiple_gpu_counts", 1230,
     "gpu:count=1;gpu:memory_mb_min=24000;gpu:memory_mb_max=24999;gpu:count=2;gpu:count=3;gpu:count=4;",
     dict(gpu=[dict(count=1, memory_mb_min=24000, memory_mb_max=24999), dict(count=2), dict(count=3), dict(count=4)],
          ram_mb=None, storage_gb=None, cpu=None)),
    ("ng_test_parsing_groups_from_string/a100", 59, "gpu:model=A100;gpu:count=8",
     dict(gpu=[dict(count=8, model="A100")], ram_mb=None, storage_gb=None, cpu=None)),
]

GPU_FIELDS = ["count", "model", "memory_mb", "memory_mb_min", "memory_mb_max", "total_memory_min", "total_memory_max"]
