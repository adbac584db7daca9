from helmpy_b200.core.helm import helm, helm_batch, non_islanding_branches
from helmpy_b200.core.classes import create_case_data_object_from_xlsx
from helmpy_b200.core.nr import nr
from helmpy_b200.core.nr import nr_ds
from helmpy_b200.core.matpower import create_case_data_object_from_matpower, matpower_to_xlsx
from helmpy_b200.core.nr import nr_batch
