/*
 * libyaml-subset compatibility implementation (from scratch; see yaml.h).
 * Reads the whole file, splits it into (indent, text) lines with comments
 * removed, and builds the event queue by recursive descent over indentation.
 */
#include "yaml.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>

struct mcr_yaml_ev { yaml_event_type_t type; char *value; };

typedef struct { int indent; char *text; } yline;

static int push_event(yaml_parser_t *p, yaml_event_type_t t, const char *v) {
    if (p->n_events == p->cap_events) {
        size_t nc = p->cap_events ? p->cap_events * 2 : 64;
        struct mcr_yaml_ev *ne = (struct mcr_yaml_ev *)realloc(p->events, nc * sizeof(*ne));
        if (!ne) return 0;
        p->events = ne; p->cap_events = nc;
    }
    p->events[p->n_events].type = t;
    p->events[p->n_events].value = v ? strdup(v) : NULL;
    p->n_events++;
    return 1;
}

/* strip a trailing comment that is outside quotes, then trailing spaces */
static void strip_comment(char *s) {
    int in_s = 0, in_d = 0;
    for (char *c = s; *c; c++) {
        if (*c == '\'' && !in_d) in_s = !in_s;
        else if (*c == '"' && !in_s) in_d = !in_d;
        else if (*c == '#' && !in_s && !in_d && (c == s || isspace((unsigned char)c[-1]))) { *c = 0; break; }
    }
    size_t n = strlen(s);
    while (n && isspace((unsigned char)s[n - 1])) s[--n] = 0;
}

/* scalar text -> value (unquote) ; returns malloc'd string */
static char *scalar_value(const char *s) {
    while (*s && isspace((unsigned char)*s)) s++;
    size_t n = strlen(s);
    while (n && isspace((unsigned char)s[n - 1])) n--;
    char *out = (char *)malloc(n + 1);
    if (!out) return NULL;
    if (n >= 2 && ((s[0] == '"' && s[n - 1] == '"') || (s[0] == '\'' && s[n - 1] == '\''))) {
        size_t o = 0;
        for (size_t i = 1; i + 1 < n; i++) {
            if (s[0] == '"' && s[i] == '\\' && i + 2 < n) {
                i++;
                out[o++] = (s[i] == 'n') ? '\n' : (s[i] == 't') ? '\t' : s[i];
            } else if (s[0] == '\'' && s[i] == '\'' && s[i + 1] == '\'' && i + 2 < n) {
                out[o++] = '\''; i++;
            } else out[o++] = s[i];
        }
        out[o] = 0;
    } else {
        memcpy(out, s, n); out[n] = 0;
    }
    return out;
}

/* find "key:" separator outside quotes; returns pointer to ':' or NULL */
static char *find_key_sep(char *s) {
    int in_s = 0, in_d = 0;
    for (char *c = s; *c; c++) {
        if (*c == '\'' && !in_d) in_s = !in_s;
        else if (*c == '"' && !in_s) in_d = !in_d;
        else if (*c == ':' && !in_s && !in_d && (c[1] == 0 || isspace((unsigned char)c[1]))) return c;
    }
    return NULL;
}

static int parse_block(yaml_parser_t *p, yline *L, size_t n, size_t *idx, int indent);

static int emit_scalar_text(yaml_parser_t *p, const char *text) {
    char *v = scalar_value(text);
    if (!v) return 0;
    int ok = push_event(p, YAML_SCALAR_EVENT, v);
    free(v);
    return ok;
}

/* value after "key:" or "- " : inline scalar, nested block, or empty scalar */
static int parse_value(yaml_parser_t *p, yline *L, size_t n, size_t *idx, const char *inline_text, int parent_indent) {
    while (*inline_text && isspace((unsigned char)*inline_text)) inline_text++;
    if (*inline_text) return emit_scalar_text(p, inline_text);
    if (*idx < n && L[*idx].indent > parent_indent) return parse_block(p, L, n, idx, L[*idx].indent);
    /* a block sequence may sit at the same indent as its parent key */
    if (*idx < n && L[*idx].indent == parent_indent && L[*idx].text[0] == '-' &&
        (L[*idx].text[1] == 0 || L[*idx].text[1] == ' '))
        return parse_block(p, L, n, idx, parent_indent);
    return push_event(p, YAML_SCALAR_EVENT, "");
}

static int parse_block(yaml_parser_t *p, yline *L, size_t n, size_t *idx, int indent) {
    if (*idx >= n) return 1;
    const int is_seq = L[*idx].text[0] == '-' && (L[*idx].text[1] == 0 || L[*idx].text[1] == ' ');
    if (!push_event(p, is_seq ? YAML_SEQUENCE_START_EVENT : YAML_MAPPING_START_EVENT, NULL)) return 0;
    while (*idx < n && L[*idx].indent == indent) {
        char *t = L[*idx].text;
        if (is_seq) {
            if (!(t[0] == '-' && (t[1] == 0 || t[1] == ' '))) break;
            (*idx)++;
            if (!parse_value(p, L, n, idx, t + 1, indent)) return 0;
        } else {
            if (t[0] == '-' && (t[1] == 0 || t[1] == ' ')) break;
            char *sep = find_key_sep(t);
            if (!sep) { p->error = 1; return 0; }
            *sep = 0;
            if (!emit_scalar_text(p, t)) return 0;
            (*idx)++;
            if (!parse_value(p, L, n, idx, sep + 1, indent)) return 0;
        }
    }
    return push_event(p, is_seq ? YAML_SEQUENCE_END_EVENT : YAML_MAPPING_END_EVENT, NULL);
}

static int build_events(yaml_parser_t *p) {
    p->built = 1;
    if (!p->input) { p->error = 1; return 0; }
    /* slurp */
    size_t cap = 4096, len = 0;
    char *buf = (char *)malloc(cap);
    if (!buf) return 0;
    for (;;) {
        if (len + 2048 > cap) { cap *= 2; char *nb = (char *)realloc(buf, cap); if (!nb) { free(buf); return 0; } buf = nb; }
        size_t r = fread(buf + len, 1, 2048, p->input);
        len += r;
        if (r < 2048) break;
    }
    buf[len] = 0;
    /* split */
    size_t nl = 0, capl = 64;
    yline *L = (yline *)malloc(capl * sizeof(yline));
    if (!L) { free(buf); return 0; }
    char *save = NULL;
    for (char *line = strtok_r(buf, "\n", &save); line; line = strtok_r(NULL, "\n", &save)) {
        size_t ll = strlen(line);
        if (ll && line[ll - 1] == '\r') line[ll - 1] = 0;
        int ind = 0;
        while (line[ind] == ' ') ind++;
        char *t = line + ind;
        strip_comment(t);
        if (!*t || strcmp(t, "---") == 0 || strcmp(t, "...") == 0) continue;
        if (nl == capl) { capl *= 2; yline *nL = (yline *)realloc(L, capl * sizeof(yline)); if (!nL) { free(L); free(buf); return 0; } L = nL; }
        L[nl].indent = ind; L[nl].text = t; nl++;
    }
    int ok = push_event(p, YAML_STREAM_START_EVENT, NULL);
    if (ok && nl) {
        ok = push_event(p, YAML_DOCUMENT_START_EVENT, NULL);
        size_t idx = 0;
        if (ok) ok = parse_block(p, L, nl, &idx, L[0].indent);
        if (ok && idx != nl) { p->error = 1; ok = 0; }
        if (ok) ok = push_event(p, YAML_DOCUMENT_END_EVENT, NULL);
    }
    if (ok) ok = push_event(p, YAML_STREAM_END_EVENT, NULL);
    free(L); free(buf);
    return ok;
}

int yaml_parser_initialize(yaml_parser_t *parser) {
    if (!parser) return 0;
    memset(parser, 0, sizeof(*parser));
    return 1;
}

void yaml_parser_delete(yaml_parser_t *parser) {
    if (!parser) return;
    for (size_t i = 0; i < parser->n_events; i++) free(parser->events[i].value);
    free(parser->events);
    memset(parser, 0, sizeof(*parser));
}

void yaml_parser_set_input_file(yaml_parser_t *parser, FILE *file) { parser->input = file; }

int yaml_parser_parse(yaml_parser_t *parser, yaml_event_t *event) {
    memset(event, 0, sizeof(*event));
    if (!parser->built && !build_events(parser)) { parser->error = 1; return 0; }
    if (parser->error) return 0;
    if (parser->next_event >= parser->n_events) { event->type = YAML_NO_EVENT; return 1; }
    struct mcr_yaml_ev *e = &parser->events[parser->next_event++];
    event->type = e->type;
    if (e->type == YAML_SCALAR_EVENT) {
        event->data.scalar.value = (yaml_char_t *)strdup(e->value ? e->value : "");
        if (!event->data.scalar.value) return 0;
        event->data.scalar.length = strlen((char *)event->data.scalar.value);
        event->data.scalar.plain_implicit = 1;
    }
    return 1;
}

void yaml_event_delete(yaml_event_t *event) {
    if (!event) return;
    if (event->type == YAML_SCALAR_EVENT) free(event->data.scalar.value);
    memset(event, 0, sizeof(*event));
}
