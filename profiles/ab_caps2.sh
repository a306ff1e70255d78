# NOTE: the cap_* / stagger / carveout options exist only in commit 33a45dc (work-item launches); results: profiles/ab_caps_r2.txt, DESIGN.md section 12
. profiles/ab_caps.sh.inc
run --opt carveout=72
run --opt carveout=72 --opt cap_round=2 --opt cap_factor=2 --opt cap_back=2 --opt stagger=1 --opt chunk=102
run --opt carveout=72 --opt cap_round=2 --opt cap_factor=1 --opt cap_back=2 --opt stagger=1 --opt chunk=102
run --opt carveout=72 --opt stagger=1 --opt chunk=102
run --opt carveout=100 --opt cap_round=2 --opt cap_factor=2 --opt cap_back=2 --opt stagger=1 --opt chunk=102
