"""Relational models shared by the golden generator (built with the reference's classes) and the
tests (built with this repo's): ``ns`` supplies LV / Atom / ParamF / RelationalGraph / Domain and
the potential classes."""


def rgm_relational(ns, n_category=5, n_bank=3):
    """The relational Gaussian model of the reference's Demo/Data/RGM/Generator.py, smaller."""
    d = ns.Domain((-50, 50), continuous=True)
    p1 = ns.GaussianPotential([0., 0.], [[10., -7.], [-7., 10.]])
    p2 = ns.GaussianPotential([0., 0.], [[10., 5.], [5., 10.]])
    p3 = ns.GaussianPotential([0., 0.], [[10., 7.], [7., 10.]])
    lv_recession = ns.LV(("all",))
    lv_category = ns.LV([f"c{i}" for i in range(n_category)])
    lv_bank = ns.LV([f"b{i}" for i in range(n_bank)])
    atoms = (ns.Atom(d, (lv_recession,), name="recession"), ns.Atom(d, (lv_bank,), name="revenue"),
             ns.Atom(d, (lv_category, lv_bank), name="loss"), ns.Atom(d, (lv_category,), name="market"))
    fs = (ns.ParamF(p1, nb=("recession($all)", "market(c)")),
          ns.ParamF(p2, nb=("market(c)", "loss(c,b)")),
          ns.ParamF(p3, nb=("loss(c,b)", "revenue(b)")))
    data = {("market", "c1"): 3.5, ("loss", "c0", "b2"): -2.0, ("loss", "c4", "b0"): -2.0, ("revenue", "b1"): 10.0}
    return ns.RelationalGraph(atoms, fs), data


def friends_relational(ns, n_person=4):
    """Hybrid friends model with a constrained parametric factor (x != y) and a constant."""
    d_bool = ns.Domain((0, 1))
    d_real = ns.Domain((-10, 10), continuous=True)
    people = ns.LV([f"p{i}" for i in range(n_person)])
    atoms = (ns.Atom(d_real, (people,), name="mood"), ns.Atom(d_bool, (people, people), name="friends"),
             ns.Atom(d_real, (people,), name="base"))
    link = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], x[2]), w=0.7)
    prior = ns.MLNPotential(lambda x: ns.eq_op(x[0], x[1]), w=0.3)
    fs = (ns.ParamF(link, nb=("friends(x,y)", "mood(x)", "mood(y)"), constrain=lambda s: s["x"] != s["y"]),
          ns.ParamF(prior, nb=("mood(x)", "base($p0)")))
    data = {("base", "p0"): 1.5}
    for i in range(n_person):
        for j in range(n_person):
            if i != j:
                data[("friends", f"p{i}", f"p{j}")] = (i + j) % 2
    return ns.RelationalGraph(atoms, fs), data


RELATIONAL = {"rgm": rgm_relational, "friends": friends_relational}
