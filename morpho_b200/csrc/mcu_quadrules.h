// mcu_quadrules.h — the quadrature rules of the reference's rule-table integrator (src/geometry/integrate.c:812-1982)
// and the rule selection of integrator_configure (integrate.c:2592-2704), host side.
//
// The reference tabulates 33 rules as arrays of nodes (barycentric) and weights.  Here they are GENERATED where a
// formula exists and stored by symmetry orbit where none does; `build()` expands them into the reference's node order,
// which matters twice: an extension rule re-uses the function values of the rule it extends (its nodes start with
// that rule's nodes, integrator_quadrature integrate.c:2445-2504), and the order of the weighted sums fixes the last
// bits of every estimate.
//   lines        midpoint/simpson (the reference's "simpson" has one node: kept as it is), Gauss-Kronrod pairs
//                1-3, 2-5, 5-11, 7-15 (classical constants);
//   triangles    tri0/4/10/20 and grundmann2d0..5: Grundmann-Moeller rules (SIAM J. Numer. Anal. 15 (1978) 282),
//                w_i = (-1)^i 2^-2s (d+n-2i)^d / (i! (d+n-i)!) * n!, nodes (2 beta + 1)/(d+n-2i), in two different
//                enumeration orders; cubtri0/7/19: Laurie's CUBTRI pair (ACM TOMS 8 (1982) 210);
//   tetrahedra   keast4, keast5 (Keast, CMAME 55 (1986) 339), tet5, tet6 (Shunn & Ham, JCAM 236 (2012) 4348),
//                grundmann3d0..5.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

struct McuQuadRule {
    std::string name;
    int grade = 0, order = 0, nnodes = 0;
    int family = -1;                 // rules of one family share a node array (index into families)
    std::vector<double> weights;
    int ext = -1;                    // index of the extension rule (same family, more nodes), or -1
};

struct McuQuadRules {
    std::vector<McuQuadRule> rules;                 // in the order of quadrules[] (integrate.c:1958-1975)
    std::vector<std::vector<double>> families;      // node arrays, (grade + 1) barycentric coordinates per node
    int default_rule[4] = {-1, -1, -1, -1};         // defaultquadrule[] (integrate.c:1977-1982): gauss5, cubtri7, grundmann3d0

    static const McuQuadRules &get() { static McuQuadRules R = build(); return R; }

    int by_name(int grade, const char *name) const {               // integrator_matchrulebyname
        for (size_t i = 0; i < rules.size(); i++) if (rules[i].grade == grade && rules[i].name == name) return (int) i;
        return -1;
    }
    int by_extension(int rule) const {                             // integrator_matchrulebyextension: first rule whose ext is `rule`
        for (size_t i = 0; i < rules.size(); i++) if (rules[i].ext == rule) return (int) i;
        return -1;
    }
    int by_order(int grade, int minorder, int maxorder, bool highest) const {      // integrator_matchrulebyorder
        int best = -1, bestorder = highest ? -1 : INT_MAX;
        for (size_t i = 0; i < rules.size(); i++) {
            const McuQuadRule &r = rules[i];
            if (r.grade != grade) continue;
            if (r.order >= minorder && r.order <= maxorder && ((highest && r.order > bestorder) || (!highest && r.order < bestorder))) { best = (int) i; bestorder = r.order; }
        }
        return best;
    }
    // integrator_configure (integrate.c:2652-2704).  Returns false when the reference would raise IntgrtrRlNtFnd /
    // IntgrtrRlUnavlb (the caller leaves such calls to the reference, which raises them).
    bool configure(int grade, bool adapt, int order, const char *name, int &rule, int &errrule) const {
        rule = errrule = -1;
        if (name && name[0]) rule = by_name(grade, name);
        else if (order >= 0) rule = by_order(grade, order, INT_MAX, false);
        else rule = (grade >= 1 && grade <= 3) ? default_rule[grade] : -1;
        if (rule < 0) return false;
        if (adapt && rules[rule].ext < 0) {
            int base = by_extension(rule);
            if (base >= 0) rule = base;
            else {
                errrule = by_order(grade, rules[rule].order + 1, INT_MAX, false);
                if (errrule < 0) errrule = by_order(grade, 0, rules[rule].order - 1, true);
                if (errrule < 0) return false;
            }
            if (errrule >= 0 && rules[rule].order > rules[errrule].order) std::swap(rule, errrule);
        }
        return true;
    }

private:
    static double fact(int k) { double f = 1.0; for (int i = 2; i <= k; i++) f *= i; return f; }
    // Grundmann-Moeller weight of the group with denominator d + n - 2i in the rule of index s (degree d = 2s + 1) on
    // the n-simplex, normalised to weights that sum to 1
    static double gm_weight(int n, int s, int i) {
        const int d = 2 * s + 1;
        return ((i & 1) ? -1.0 : 1.0) * std::ldexp(1.0, -2 * s) * std::pow((double) (d + n - 2 * i), d) / (fact(i) * fact(d + n - i)) * fact(n);
    }
    int add_family(std::vector<double> nodes) { families.push_back(std::move(nodes)); return (int) families.size() - 1; }
    int add_rule(const char *name, int grade, int order, int nnodes, int family, std::vector<double> w) {
        McuQuadRule r;
        r.name = name; r.grade = grade; r.order = order; r.nnodes = nnodes; r.family = family; r.weights = std::move(w);
        rules.push_back(r);
        return (int) rules.size() - 1;
    }
    // distinct permutations of a multiset in ascending lexicographic order
    static void lex_perms(std::vector<double> v, std::vector<double> &out) {
        std::sort(v.begin(), v.end());
        do out.insert(out.end(), v.begin(), v.end()); while (std::next_permutation(v.begin(), v.end()));
    }
    // orbit of a generator written as a pattern string ("abbb babb ..."), values for a, b, c
    static void pattern(const char *perms, double a, double b, double c, std::vector<double> &out) {
        for (const char *q = perms; *q;) {
            int len = 0;
            while (q[len] && q[len] != ' ') { out.push_back(q[len] == 'a' ? a : (q[len] == 'b' ? b : c)); len++; }
            q += len;
            while (*q == ' ') q++;
        }
    }

    static McuQuadRules build() {
        McuQuadRules R;
        // ---- lines: nodes are (t, 1 - t) --------------------------------------------------------------------
        auto line_nodes = [](std::initializer_list<double> t) { std::vector<double> n; for (double x : t) { n.push_back(x); n.push_back(1.0 - x); } return n; };
        {   // midpoint with the reference's one-node "simpson" extension (integrate.c:813-845)
            int f = R.add_family({0.5, 0.5, 0.0, 1.0, 1.0, 0.0});
            int mid = R.add_rule("midpoint", 1, 1, 1, f, {1.0});
            int sim = R.add_rule("simpson", 1, 3, 1, f, {0.66666666666666667, 0.16666666666666667, 0.16666666666666667});
            R.rules[mid].ext = sim;
        }
        {   // Gauss 1 / Kronrod 3
            int f = R.add_family(line_nodes({0.5, 0.11270166537925831148, 0.88729833462074168852}));
            int g = R.add_rule("gauss1", 1, 1, 1, f, {1.0});
            int k = R.add_rule("kronrod3", 1, 3, 3, f, {0.4444444444444444444445, 0.2777777777777777777778, 0.277777777777777777778});
            R.rules[g].ext = k;
        }
        {   // Gauss 2 / Kronrod 5
            int f = R.add_family(line_nodes({0.21132486540518711775, 0.78867513459481288225, 0.037089950113724269217, 0.5, 0.96291004988627573078}));
            int g = R.add_rule("gauss2", 1, 3, 2, f, {0.5, 0.5});
            int k = R.add_rule("kronrod5", 1, 7, 5, f, {0.2454545454545454545455, 0.245454545454545454546, 0.098989898989898989899, 0.3111111111111111111111, 0.098989898989898989899});
            R.rules[g].ext = k;
        }
        {   // Gauss 5 / Kronrod 11
            int f = R.add_family(line_nodes({0.046910077030668003601, 0.23076534494715845448, 0.5, 0.76923465505284154552, 0.9530899229693319964,
                                             0.0079573199525787677519, 0.12291663671457538978, 0.36018479341910840329, 0.63981520658089159671,
                                             0.87708336328542461022, 0.99204268004742123225}));
            int g = R.add_rule("gauss5", 1, 9, 5, f, {0.118463442528094543757, 0.2393143352496832340206, 0.2844444444444444444444, 0.2393143352496832340206, 0.118463442528094543757});
            int k = R.add_rule("kronrod11", 1, 16, 11, f, {0.0576166583112366970123, 0.12052016961432379335, 0.1414937089287456066021, 0.12052016961432379335,
                                                           0.0576166583112366970123, 0.02129101837554091643225, 0.093400398278246328734, 0.136424900956279461171,
                                                           0.136424900956279461171, 0.0934003982782463287339, 0.0212910183755409164322});
            R.rules[g].ext = k;
        }
        {   // Gauss 7 / Kronrod 15
            int f = R.add_family(line_nodes({0.0254460438286207377369, 0.1292344072003027800681, 0.2970774243113014165467, 0.5, 0.7029225756886985834533,
                                             0.8707655927996972199320, 0.9745539561713792622631,
                                             0.0042723144395936803966, 0.0675677883201154636052, 0.2069563822661544348530, 0.3961075224960507661997,
                                             0.6038924775039492338004, 0.7930436177338455651471, 0.9324322116798845363949, 0.9957276855604063196035}));
            int g = R.add_rule("gauss7", 1, 13, 7, f, {0.0647424830844348466355, 0.1398526957446383339505, 0.1909150252525594724752, 0.20897959183673469387755,
                                                       0.1909150252525594724752, 0.13985269574463833395075, 0.0647424830844348466353});
            int k = R.add_rule("kronrod15", 1, 22, 15, f, {0.0315460463149892766454, 0.070326629857762959373, 0.0951752890323927049567, 0.104741070542363914007,
                                                           0.095175289032392704957, 0.0703266298577629593726, 0.0315460463149892766454,
                                                           0.0114676610052646124819, 0.0523950051611250919200, 0.0845023633196339514133, 0.1022164700376494462071,
                                                           0.1022164700376494462071, 0.0845023633196339514133, 0.05239500516112509192, 0.01146766100526461248187});
            R.rules[g].ext = k;
        }
        // ---- triangles ----------------------------------------------------------------------------------------
        auto gm_weights = [&](int n, int s, const std::vector<int> &group_sizes) {      // groups in ascending denominator
            std::vector<double> w;
            for (int g = 0; g <= s; g++) w.insert(w.end(), (size_t) group_sizes[g], gm_weight(n, s, s - g));
            return w;
        };
        {   // "tri" family (Walkington's order, integrate.c:1068-1093): centroid; (3,1,1)/5; (5,1,1)/7, (3,3,1)/7; (7,1,1)/9, (3,5,1)/9.., (3,3,3)/9
            std::vector<double> n;
            auto node = [&](double a, double b, double c, double den) { n.push_back(a / den); n.push_back(b / den); n.push_back(c / den); };
            node(1, 1, 1, 3);
            node(3, 1, 1, 5); node(1, 3, 1, 5); node(1, 1, 3, 5);
            node(5, 1, 1, 7); node(1, 5, 1, 7); node(1, 1, 5, 7); node(3, 3, 1, 7); node(3, 1, 3, 7); node(1, 3, 3, 7);
            node(7, 1, 1, 9); node(1, 7, 1, 9); node(1, 1, 7, 9);
            node(3, 5, 1, 9); node(3, 1, 5, 9); node(5, 3, 1, 9); node(5, 1, 3, 9); node(1, 3, 5, 9); node(1, 5, 3, 9);
            node(3, 3, 3, 9);
            int f = R.add_family(n);
            const std::vector<int> sizes = {1, 3, 6, 10};
            const int orders[4] = {1, 3, 5, 8}, counts[4] = {1, 4, 10, 20};
            const char *names[4] = {"tri0", "tri4", "tri10", "tri20"};
            int prev = -1;
            for (int s = 0; s < 4; s++) {
                int r = R.add_rule(names[s], 2, orders[s], counts[s], f, gm_weights(2, s, sizes));
                if (prev >= 0) R.rules[prev].ext = r;
                prev = r;
            }
        }
        {   // CUBTRI: Radon's 7-point rule and Laurie's 19-point extension
            std::vector<double> n;
            const char *P3 = "abb bba bab";
            n.insert(n.end(), {0.333333333333333333, 0.333333333333333333, 0.333333333333333333});
            pattern(P3, 0.797426985353087322, 0.101286507323456339, 0, n);
            pattern(P3, 0.0597158717897698205, 0.47014206410511509, 0, n);
            pattern(P3, 0.941038278231120867, 0.0294808608844395667, 0, n);
            pattern(P3, 0.535795346449899265, 0.232102326775050368, 0, n);
            pattern("abc bac cab cba acb bca", 0.0294808608844395667, 0.232102326775050368, 0.738416812340510066, n);
            int f = R.add_family(n);
            int c0 = R.add_rule("cubtri0", 2, 1, 1, f, {1.0});
            std::vector<double> w7 = {0.225000000000000000};
            w7.insert(w7.end(), 3, 0.125939180544827153); w7.insert(w7.end(), 3, 0.132394152788506181);
            int c7 = R.add_rule("cubtri7", 2, 5, 7, f, w7);
            std::vector<double> w19 = {0.0378610912003146833};
            w19.insert(w19.end(), 3, 0.0376204254131829721); w19.insert(w19.end(), 3, 0.0783573522441173376);
            w19.insert(w19.end(), 3, 0.0134442673751654019); w19.insert(w19.end(), 3, 0.116271479656965896);
            w19.insert(w19.end(), 6, 0.0375097224552317488);
            int c19 = R.add_rule("cubtri19", 2, 8, 19, f, w19);
            R.rules[c0].ext = c7; R.rules[c7].ext = c19;
        }
        {   // grundmann2d0..5 (Burkardt's enumeration, integrate.c:1237-1300): per denominator 2s+3 the points
            // (2 b0 + 1, 2 b1 + 1, 2 b2 + 1) with b1 descending from s and b0 descending inside
            std::vector<double> n;
            std::vector<int> sizes;
            for (int s = 0; s <= 5; s++) {
                const double den = 2 * s + 3;
                int cnt = 0;
                for (int b1 = s; b1 >= 0; b1--)
                    for (int b0 = s - b1; b0 >= 0; b0--) {
                        const int b2 = s - b1 - b0;
                        n.push_back((2 * b0 + 1) / den); n.push_back((2 * b1 + 1) / den); n.push_back((2 * b2 + 1) / den);
                        cnt++;
                    }
                sizes.push_back(cnt);
            }
            int f = R.add_family(n);
            int prev = -1, total = 0;
            for (int s = 0; s <= 5; s++) {
                total += sizes[s];
                char name[24];
                snprintf(name, sizeof(name), "grundmann2d%d", s);
                int r = R.add_rule(name, 2, 2 * s + 1, total, f, gm_weights(2, s, sizes));
                if (prev >= 0) R.rules[prev].ext = r;
                prev = r;
            }
        }
        // ---- tetrahedra -----------------------------------------------------------------------------------------
        const char *P4 = "abbb babb bbab bbba", *P6 = "aacc acac acca caac caca ccaa";
        const char *P12A = "abcc bacc acbc bcac accb bcca cabc cbac cacb cbca ccab ccba";
        const char *P12B = "abbc babc bbac abcb bacb bbca acbb bcab bcba cabb cbab cbba";
        {   // Keast's 11-point rule of degree 4
            std::vector<double> n = {0.25, 0.25, 0.25, 0.25};
            pattern(P4, 0.78571428571428571, 0.071428571428571428, 0, n);
            pattern(P6, 0.39940357616679922, 0, 0.10059642383320078, n);
            std::vector<double> w = {-0.07893333333333333};
            w.insert(w.end(), 4, 0.04573333333333333333); w.insert(w.end(), 6, 0.149333333333333328);
            R.add_rule("keast4", 3, 4, 11, R.add_family(n), w);
        }
        {   // Keast's 15-point rule of degree 5
            std::vector<double> n = {0.25, 0.25, 0.25, 0.25};
            pattern(P4, 0.0, 0.3333333333333333, 0, n);
            pattern(P4, 0.72727272727272727, 0.090909090909090909, 0, n);
            pattern(P6, 0.066550153573664281, 0, 0.43344984642633573, n);
            std::vector<double> w = {0.181702068582535114};
            w.insert(w.end(), 4, 0.0361607142857142958); w.insert(w.end(), 4, 0.069871494516173845); w.insert(w.end(), 6, 0.06569484936831872);
            R.add_rule("keast5", 3, 5, 15, R.add_family(n), w);
        }
        {   // Shunn & Ham, 35 points
            std::vector<double> n, w;
            pattern(P4, 0.9197896733368800, 0.0267367755543735, 0, n); w.insert(w.end(), 4, 0.0021900463965388);
            pattern(P12A, 0.1740356302468940, 0.7477598884818090, 0.0391022406356488, n); w.insert(w.end(), 12, 0.0143395670177665);
            pattern(P6, 0.4547545999844830, 0, 0.0452454000155172, n); w.insert(w.end(), 6, 0.0250305395686746);
            pattern(P12B, 0.5031186450145980, 0.2232010379623150, 0.0504792790607720, n); w.insert(w.end(), 12, 0.0479839333057554);
            n.insert(n.end(), 4, 0.25); w.push_back(0.0931745731195340);
            R.add_rule("tet5", 3, 7, 35, R.add_family(n), w);
        }
        {   // Shunn & Ham, 56 points
            std::vector<double> n, w;
            pattern(P4, 0.9551438045408220, 0.0149520651530592, 0, n); w.insert(w.end(), 4, 0.0010373112336140);
            pattern(P12A, 0.7799760084415400, 0.1518319491659370, 0.0340960211962615, n); w.insert(w.end(), 12, 0.0096016645399480);
            pattern(P12A, 0.3549340560639790, 0.5526556431060170, 0.0462051504150017, n); w.insert(w.end(), 12, 0.0164493976798232);
            pattern(P12B, 0.5381043228880020, 0.2281904610687610, 0.0055147549744775, n); w.insert(w.end(), 12, 0.0153747766513310);
            pattern(P12B, 0.1961837595745600, 0.3523052600879940, 0.0992057202494530, n); w.insert(w.end(), 12, 0.0293520118375230);
            pattern(P4, 0.5965649956210170, 0.1344783347929940, 0, n); w.insert(w.end(), 4, 0.0366291366405108);
            R.add_rule("tet6", 3, 9, 56, R.add_family(n), w);
        }
        {   // grundmann3d0..5 (integrate.c:1641-1800): per denominator 2s+4 the partitions of it into four odd parts in
            // descending order of their ascending-sorted tuple, every orbit's distinct permutations in ascending order
            std::vector<double> n;
            std::vector<int> sizes;
            for (int s = 0; s <= 5; s++) {
                const int den = 2 * s + 4;
                std::vector<std::vector<int>> parts;
                for (int a = 1; a <= den; a += 2) for (int b = a; b <= den; b += 2) for (int c = b; c <= den; c += 2) {
                    int d = den - a - b - c;
                    if (d >= c && (d & 1)) parts.push_back({a, b, c, d});
                }
                std::sort(parts.begin(), parts.end(), [](const std::vector<int> &x, const std::vector<int> &y) { return x > y; });
                size_t before = n.size();
                for (auto &p : parts) {
                    std::vector<double> v = {p[0] / (double) den, p[1] / (double) den, p[2] / (double) den, p[3] / (double) den};
                    lex_perms(v, n);
                }
                sizes.push_back((int) ((n.size() - before) / 4));
            }
            int f = R.add_family(n);
            int prev = -1, total = 0, first = -1;
            for (int s = 0; s <= 5; s++) {
                total += sizes[s];
                char name[24];
                snprintf(name, sizeof(name), "grundmann3d%d", s);
                int r = R.add_rule(name, 3, 2 * s + 1, total, f, gm_weights(3, s, sizes));
                if (prev >= 0) R.rules[prev].ext = r;
                if (s == 0) first = r;
                prev = r;
            }
            R.default_rule[3] = first;
        }
        R.default_rule[1] = R.by_name(1, "gauss5");
        R.default_rule[2] = R.by_name(2, "cubtri7");
        return R;
    }
};
